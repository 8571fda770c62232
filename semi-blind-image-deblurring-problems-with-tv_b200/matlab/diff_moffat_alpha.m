function H = diff_moffat_alpha(im_shape, size, a, b)
% Drop-in for utils/diff_moffat_alpha.m:1-22 (PSF / derivative spectrum = resize(kernel, im_shape)).
H = sbd_mex('spectrum', double(im_shape(1:2)), 1, size, 0, [a b], 1);
