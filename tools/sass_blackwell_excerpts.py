import re,subprocess
lib='semi-blind-image-deblurring-problems-with-tv_b200/lib/libsbd.so'
out=subprocess.run(["cuobjdump","-sass",lib],capture_output=True,text=True).stdout
md=["# Blackwell asynchronous-copy instructions in libsbd.so (r02) - `cuobjdump -sass` excerpts","",
"The v2 FFT passes (csrc/fft2.cuh) move their data with the bulk / tensor copy engines. PTX -> SASS: `cp.async.bulk` -> `UBLKCP`,",
"`cp.async.bulk.tensor.3d` -> `UTMALDG.3D` / `UTMASTG.3D`, `mbarrier.*` -> `SYNCS.*`, `fence.proxy.async` -> `FENCE.VIEW.ASYNC`.","",
"Regenerate with `python tools/sass_blackwell_excerpts.py` (no GPU needed).",""]
for fn in re.split(r"\n\s*Function : ",out)[1:]:
    name=fn.split("\n",1)[0].strip()
    if not any(k in name for k in ("k_cols2ILi4096","k_rows2_fwdILi4096","k_rows2_invILi4096")): continue
    lines=[l.rstrip() for l in fn.split("\n") if re.search(r"UBLKCP|UTMA|SYNCS|FENCE\.VIEW|UTMACMDFLUSH|UTMAPF",l)]
    tot=len(re.findall(r"/\*[0-9a-f]{4,}\*/\s+\S",fn))
    dem=subprocess.run(["c++filt",name],capture_output=True,text=True).stdout.strip()
    md+= [f"## `{dem}`","",f"({tot//2 if False else tot} SASS lines)","","```"]+[re.sub(r"\s+/\* 0x[0-9a-f]+ \*/","",l) for l in lines]+["```",""]
open('profiles/r02_sass_blackwell.md','w').write("\n".join(md))
