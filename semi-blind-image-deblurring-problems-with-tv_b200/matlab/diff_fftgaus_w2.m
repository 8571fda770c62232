function H = diff_fftgaus_w2(im_shape, taille, w1, w2, phi)
% Drop-in for utils/diff_fftgaus_w2.m:2-26 (PSF / derivative spectrum = resize(kernel, im_shape)).
H = sbd_mex('spectrum', double(im_shape(1:2)), 0, taille, phi, [w1 w2], 2);
