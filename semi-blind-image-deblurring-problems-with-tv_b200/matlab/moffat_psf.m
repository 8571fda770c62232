function H = moffat_psf(im_shape, size, a, b)
% Drop-in for utils/moffat_psf.m:2-23 (PSF / derivative spectrum = resize(kernel, im_shape)).
H = sbd_mex('spectrum', double(im_shape(1:2)), 1, size, 0, [a b], 0);
