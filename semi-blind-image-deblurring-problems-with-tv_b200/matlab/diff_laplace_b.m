function H = diff_laplace_b(im_shape, size, b)
% Drop-in for utils/diff_laplace_b.m:1-19 (PSF / derivative spectrum = resize(kernel, im_shape)).
H = sbd_mex('spectrum', double(im_shape(1:2)), 2, size, 0, b, 1);
