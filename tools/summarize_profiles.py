#!/usr/bin/env python
"""Turn the ncu CSV exports of one bench run into the tracked summaries under profiles/.

    python tools/summarize_profiles.py gpurun_out/launches_r01g.csv gpurun_out/step_r01g_raw.csv [round tag, default r01]

  launches csv : ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file <csv> python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-size-sweep
  raw csv      : ncu --set full --clock-control none -s 46 -c 23 -o rep python bench.py ... ; ncu -i rep.ncu-rep --page raw --csv > <csv>
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def clean(n):
    n = n.replace("(int)", "").replace("(bool)", "").replace("void ", "").replace("sbd::", "")
    return re.sub(r"\(.*", "", n)


def launch_list(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    L = []
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        t = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
        t = t / 1000 if u == "ns" else (t * 1000 if u == "ms" else t)
        L.append((clean(r[ix["Kernel Name"]]), r[ix["Grid Size"]], t))
    st = [i for i, (n, _, _) in enumerate(L) if n.startswith("k_cols<4096, 2") or n.startswith("k_cols2<4096, 2")]
    a, b = st[0], st[1]
    tot = sum(t for _, _, t in L[a:b])
    agg = {}
    for n, g, t in L[a:b]:
        agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += t
    md = [f"# ncu launch list, ONE main-loop step (4096^2, 8 chains, K = 25) - {tag}, final kernels", "",
          "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 70 --csv python bench.py --steps 3 --warmup 3 --chains-per-gpu 8 --no-cpu-baseline --no-size-sweep --no-extras`",
          "(times under ncu are serialised and cold-cache: compare SHARES with the live CUDA-event phases of bench.py)", "",
          "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for n, (c, t) in agg.items():
        md.append(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    md.append(f"| **step** | {b - a} | {tot:.1f} | 100% |")
    md += ["", "Template arguments: `k_chamb_multi<T, PIPE, MINB, ZERO, EMIT>`, `k_cols2<N, MODE>` / `k_cols<N, MODE, SYM>` (MODE 1 = forward + "
           "likelihood sums, 2 = gradient multiply + inverse).", "", "Per-launch sequence:", "", "```"]
    for n, g, t in L[a:b]:
        md.append(f"{n:38s} {g:>16s} {t:10.1f} us")
    md.append("```")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches_one_step.md"), "w").write("\n".join(md) + "\n")
    return [n for n, _, t in L[a:b] if t > 50.0]        # the no-op redo launches (a few us) move no data


def full_set(path, tag, step_kernels=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]; units = rows[1]
    col = hdr.index
    sel = {}
    for r in rows[2:]:
        n = clean(r[col("Kernel Name")]); d = float(r[col("gpu__time_duration.sum")])
        if n not in sel or d > float(sel[n][col("gpu__time_duration.sum")]):
            sel[n] = r
    want = [("duration", "gpu__time_duration.sum"), ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
            ("dram throughput % of ncu peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            ("fp64 pipe active %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
            ("registers/thread", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
            ("warp instructions", "smsp__inst_executed.sum"),
            ("LSU data-pipe wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            ("L2 hit rate %", "lts__t_sector_hit_rate.pct")]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = [f"# ncu `--set full` summaries - {tag}, final kernels (B200, 4096^2, 8 chains, K = 25)", "",
           "Captured with `ncu --set full --clock-control none -s 100 -c 30 python bench.py --steps 3 --warmup 3 --chains-per-gpu 8 --no-cpu-baseline --no-size-sweep --no-extras`",
           "(one whole main-loop step) after the same command had exited 0 without ncu; exported on the GPU box with `ncu -i ... --page raw --csv`",
           "and summarised by `tools/summarize_profiles.py`. For every kernel the longest launch of the step is shown. Stall figures are warps",
           "stalled per issue-active cycle (`smsp__average_warps_issue_stalled_*_per_issue_active`).", ""]
    traffic = {}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def ms(r):
        v = float(r[col("gpu__time_duration.sum")]); u = units[col("gpu__time_duration.sum")]
        return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
    for n in sorted(sel, key=lambda n: -ms(sel[n])):
        r = sel[n]
        if ms(r) < 0.05:
            continue
        out += [f"### `{n}`", "", "| metric | value |", "|---|---|"]
        for lab, h in want:
            if h in hdr:
                out.append(f"| {lab} | {r[col(h)]} {units[col(h)]} |")
        ss = sorted(((float(r[hdr.index(h)] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stall), reverse=True)[:6]
        out.append("| top stalls (warps per issue-active cycle) | " + ", ".join(f"{k} {v:.2f}" for v, k in ss) + " |")
        out.append("")
        traffic[n] = sum(float(r[col(k)]) * mult.get(units[col(k)], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
    # merge into profiles/roofline_traffic.json (the file bench.py reads), keeping its other keys
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    cur = json.load(open(tp)) if os.path.exists(tp) else {}
    cur[f"per_kernel_{tag}_4096x8"] = traffic
    if step_kernels:
        # DRAM bytes of one whole step = sum over its launches (no-op redo launches and tiny kernels count as 0)
        tot = sum(traffic.get(n, 0.0) for n in step_kernels)
        cur["step_bytes_per_pixel_chain"] = tot / (4096.0 * 4096.0 * 8)
        cur["step_bytes_note"] = f"{tag}: sum of the per-launch DRAM bytes over the {len(step_kernels)} launches of one step, / (4096^2 x 8)"
    json.dump(cur, open(tp, "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[3] if len(sys.argv) > 3 else "r01"
    names = launch_list(sys.argv[1], tag)
    # the redo launches of the fused Chambolle kernel are no-ops (a few us): count a kernel's traffic once per REAL launch
    full_set(sys.argv[2], tag, names)
