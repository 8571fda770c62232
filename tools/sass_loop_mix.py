#!/usr/bin/env python
"""Instruction mix of the loops of one kernel, from `cuobjdump -sass`.

    python tools/sass_loop_mix.py <lib.so> <mangled-kernel-name-substring> [--dump N]

Finds every backward branch (loop back edge), and for the loop body [target, branch] prints the number of
instructions by class: fp64 (DFMA/DADD/DMUL/DSETP/...), MUFU, memory (LDG/STG/LDS/STS/CCTL/...), shuffles,
register moves (MOV/IMAD.MOV), integer/other.  --dump N writes the body of the N-th largest loop.
This is how the instruction counts quoted in DESIGN.md / profiles/ are obtained (reproducible, no GPU needed)."""
import re
import subprocess
import sys
from collections import Counter


def classify(op):
    base = op.split(".")[0]
    if base in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
        return "fp64"
    if base == "MUFU":
        return "mufu"
    if base in ("LDG", "STG", "LDS", "STS", "LD", "ST", "LDC", "LDCU", "CCTL", "ATOMG", "RED", "LDSM", "UBLKCP", "UBLKPF", "SYNCS"):
        return "mem"
    if base in ("SHFL",):
        return "shfl"
    if base == "MOV" or op.startswith("IMAD.MOV") or base in ("UMOV",):
        return "mov"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "BAR", "NOP", "BRX", "JMP"):
        return "ctrl"
    return "int_other"


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    dump = int(sys.argv[sys.argv.index("--dump") + 1]) if "--dump" in sys.argv else None
    elf = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    # split per function
    funcs = re.split(r"\n\s*Function : ", elf)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        ins = []
        for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*([^;]*);", f):
            ins.append((int(m.group(1), 16), m.group(3), m.group(4), bool(m.group(2))))
        print(f"== {name}: {len(ins)} instructions")
        loops = []
        for addr, op, args, _ in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", args)
                if t and int(t.group(1), 16) <= addr:
                    loops.append((int(t.group(1), 16), addr))
        loops.sort(key=lambda l: l[0] - l[1])
        for li, (lo, hi) in enumerate(loops):
            body = [(a, o, g) for a, o, g, _ in ins if lo <= a <= hi]
            c = Counter(classify(o) for _, o, _ in body)
            tot = len(body)
            print(f"  loop {li}: 0x{lo:x}..0x{hi:x}  {tot} instr  " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
            if dump == li:
                for a, o, g in body:
                    print(f"      /*{a:05x}*/ {o} {g}")


if __name__ == "__main__":
    main()
