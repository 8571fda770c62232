function H = laplace_psf(im_shape, size, b)
% Drop-in for utils/laplace_psf.m:1-15 (PSF / derivative spectrum = resize(kernel, im_shape)).
H = sbd_mex('spectrum', double(im_shape(1:2)), 2, size, 0, b, 0);
