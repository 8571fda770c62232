function y = TVnorm(x)
% Drop-in for utils/TVnorm.m:1-2 (isotropic TV, periodic backward differences) on the GPU.
y = sbd_mex('tvnorm', double(x));
