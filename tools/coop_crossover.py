#!/usr/bin/env python
"""chain-steps/s of the SAPG main loop, cooperative prox (tv_coop.cuh) vs fused marching kernel, over sizes and chain
counts: the measurement behind the automatic threshold of `chamb_coop`.   python tools/coop_crossover.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

cases = [(128, 1), (128, 8), (256, 1), (256, 8), (256, 32), (512, 1), (512, 8), (512, 32), (1024, 1), (1024, 2), (1024, 8),
         (2048, 1), (2048, 2), (2048, 8)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for n, ch in cases:
    row = {}
    for mode in (0, 1):
        os.environ["SBD_CHAMB_COOP"] = str(mode)
        try:
            row[mode] = bench.quick_rate(n, ch, 0, steps=40 if n <= 1024 else 12, warmup=4)
        except Exception as e:
            row[mode] = float("nan"); print("error", n, ch, mode, repr(e)[:200])
    print(f"{n:5d}^2 x {ch:3d} chains: fused {row[0]:10.1f}  coop {row[1]:10.1f} chain-steps/s  ratio {row[1]/row[0]:.2f}  ({1e6*ch/row[1]:.0f} us/step coop)", flush=True)
