function sol = diffh(x)
% Drop-in for SALSA/diffh.m:1-3.
sol = sbd_mex('diff', double(x), 1);
