#!/usr/bin/env python
"""chain-steps/s of the SAPG main loop over the segment length of the fused Chambolle kernel at mid sizes (8 chains):
the measurement behind the segment heuristic of set_geometry (sbd.cu).   python tools/seg_sweep.py [NxCH ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

cases = [(1024, 8), (2048, 8), (1024, 32), (2048, 2)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
os.environ["SBD_CHAMB_COOP"] = "0"
for n, ch in cases:
    row = []
    for seg in (0, 16, 32, 64, 128, 256):
        if seg: os.environ["SBD_CHAMB_SEG"] = str(seg)
        else: os.environ.pop("SBD_CHAMB_SEG", None)
        row.append((seg, bench.quick_rate(n, ch, 0, steps=30 if n <= 1024 else 12, warmup=4)))
    print(f"{n}^2 x {ch}: " + "  ".join(f"{'auto' if s == 0 else s}: {r:.0f}" for s, r in row), flush=True)
