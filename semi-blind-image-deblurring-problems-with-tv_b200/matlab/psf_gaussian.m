function kernel = psf_gaussian(taille, w1, w2, phi)
% Drop-in for utils/psf_gaussian.m:2-19.
kernel = sbd_mex('psf', 0, taille, phi, [w1 w2], 0);
