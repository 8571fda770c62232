function build_mex()
% BUILD_MEX  compile the gateway against libsbd.so (run once, from this directory).
here = fileparts(mfilename('fullpath'));
inc  = fullfile(here, '..', '..', 'include');
lib  = fullfile(here, '..', 'lib');
src  = fullfile(here, '..', 'mex', 'sbd_mex.c');
if exist('OCTAVE_VERSION', 'builtin')
    mkoctfile('--mex', ['-I' inc], src, ['-L' lib], '-lsbd', ['-Wl,-rpath,' lib], '-o', fullfile(here, 'sbd_mex'));
else
    mex(['-I' inc], src, ['-L' lib], '-lsbd', ['LDFLAGS=$LDFLAGS -Wl,-rpath,' lib], '-outdir', here);
end
end
