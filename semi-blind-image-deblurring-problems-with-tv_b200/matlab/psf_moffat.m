function kernel = psf_moffat(size, a, b)
% Drop-in for utils/psf_moffat.m:2-20.
kernel = sbd_mex('psf', 1, size, 0, [a b], 0);
