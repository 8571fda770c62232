function kernel = Gaussian_psf(taille, w1, w2, phi)
% Drop-in for utils/Gaussian_psf.m:2-19.
kernel = sbd_mex('psf', 0, taille, phi, [w1 w2], 0);
