function [val, iters] = sbd_max_eigenval(model, im_size, psf_size, phi, psi, tol, max_iter, x0)
% GPU version of utils/max_eigenval_Gaussian_Moffat.m:1-27 / max_eigenval_Laplace.m:28-55
% (power iteration on A'A).  x0 = [] draws the start vector on the device.
%   evMax = sbd_max_eigenval(0, im_size, psf_size, phi, [1 1], 1e-4, 1e4, randn(im_size));   % run_Gaussian_demo.m:142
if nargin < 8, x0 = []; end
[val, iters] = sbd_mex('max_eigenval', double(im_size(1:2)), model, psf_size, phi, psi, tol, max_iter, x0);
end
