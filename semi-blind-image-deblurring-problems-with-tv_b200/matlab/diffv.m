function sol = diffv(x)
% Drop-in for SALSA/diffv.m:1-3.
sol = sbd_mex('diff', double(x), 0);
