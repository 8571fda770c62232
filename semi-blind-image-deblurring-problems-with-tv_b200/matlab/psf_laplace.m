function lap = psf_laplace(size, b)
% Drop-in for utils/psf_laplace.m:1-13.
lap = sbd_mex('psf', 2, size, 0, b, 0);
