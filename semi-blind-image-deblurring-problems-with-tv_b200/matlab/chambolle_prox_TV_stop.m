function [f, px, py] = chambolle_prox_TV_stop(g, varargin)
% Drop-in for utils/chambolle_prox_TV_stop.m:1-166 - same options (case-insensitive), same
% defaults (:77-81), same failure when 'maxiter' is omitted (:80,:131), same 'dualvars' = [px py].
if (nargin - length(varargin)) ~= 1
    error('Wrong number of required parameters');
end
tau = 0.249; tol = 1e-3; lambda = 1; px0 = []; py0 = [];
for i = 1:2:(length(varargin) - 1)
    switch upper(varargin{i})
        case 'LAMBDA',  lambda  = varargin{i+1};
        case 'VERBOSE'
        case 'TOL',     tol     = varargin{i+1};
        case 'MAXITER', MaxIter = varargin{i+1};
        case 'TAU',     tau     = varargin{i+1};
        case 'DUALVARS'
            [M, N] = size(g);
            [Maux, Naux] = size(varargin{i+1});
            if M ~= Maux || Naux ~= 2 * N
                error('Wrong size of the dual variables');
            end
            px0 = varargin{i+1};
            py0 = px0(:, M+1:end);
            px0 = px0(:, 1:M);
    end
end
[f, px, py] = sbd_mex('tvprox', double(g), lambda, MaxIter, tol, tau, px0, py0);   % MaxIter undefined -> same error as the reference
