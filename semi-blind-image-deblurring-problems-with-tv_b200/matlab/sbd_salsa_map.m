function [xMAP, objective, distance, mses, n_outer] = sbd_salsa_map(y, model, psf_size, phi, psi_EB, tau, mu, maxiter, tolA, tviters, x_true)
% GPU version of the post-SAPG MAP step: SALSA/SALSA_v2.m:156-494 exactly as the demos call it
% (run_Gaussian_demo.m:229-242: 'TVINITIALIZATION',1,'TViters',10,'LS',invLS,'StopCriterion',1):
%   xMAP = sbd_salsa_map(y, 0, psf_size, phi, [w1_EB w2_EB], theta_EB*sigma_EB, theta_EB/10, 500, 1e-5, 10, x);
if nargin < 11, x_true = []; end
[xMAP, objective, distance, mses, n_outer] = sbd_mex('salsa', double(y), model, psf_size, phi, psi_EB, tau, mu, maxiter, tolA, tviters, x_true);
objective = objective(1:n_outer+1); distance = distance(1:n_outer); mses = mses(1:n_outer+1);
end
