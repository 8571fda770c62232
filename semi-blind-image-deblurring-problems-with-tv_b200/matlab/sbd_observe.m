function [y, sigma, ax_norm] = sbd_observe(x, model, psf_size, phi, psi, bsnr, noise)
% GPU version of the observation synthesis of the demos (run_Gaussian_demo.m:145-168):
% Ax = A(x; psi), sigma = norm(Ax-mean(mean(Ax)),'fro')/sqrt(numel(x)*10^(bsnr/10)), y = Ax + sigma*noise.
if nargin < 7, noise = []; end
[y, sigma, ax_norm] = sbd_mex('observe', double(x), model, psf_size, phi, psi, bsnr, noise);
end
