function [A, AT, dif] = sbd_closures(model, im_size, psf_size, phi)
% GPU versions of the closures the demo scripts build
% (run_Gaussian_demo.m:136-139, run_moffat_demo.m:134-137, run_laplace_demo.m:105-107).
%   model: 0 Gaussian (psi = w1,w2), 1 Moffat (alpha,beta), 2 Laplace (b)
A   = @(x, varargin) sbd_mex('blur', x, model, psf_size, phi, [varargin{:}], 0);
AT  = @(x, varargin) sbd_mex('blur', x, model, psf_size, phi, [varargin{:}], 1);
dif = {@(x, varargin) sbd_mex('blur', x, model, psf_size, phi, [varargin{:}], 2), ...
       @(x, varargin) sbd_mex('blur', x, model, psf_size, phi, [varargin{:}], 3)};
end
