#!/usr/bin/env python
"""SASS opcode histogram of the sampling kernels in the built library (cuobjdump; runs without a GPU) -> markdown.

The mnemonics that matter for this path: UBLKCP (TMA-engine bulk copy of the record stream), SYNCS (mbarrier),
IDP.4A (dp4a row addressing on the FMA pipe), LDS (rows, records, decode tables), LOP3 (the GF(2) algebra), VOTE (fused
input transpose), ATOMS (ring release counters, drawn bits).  No UTC*MMA / UTMALDG: the path is 1-D streams and bit work.
usage: python tools/sass_histogram.py [kernel-name regex ...]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tsim_b200", "libtsim_b200.so")
WANT = sys.argv[1:] or [r"sample_sliced_kernelILi4ELb0ELb0ELb0E", r"sample_sliced_kernelILi8ELb1ELb0ELb0E", r"sample_sliced_kernelILi4ELb0ELb1ELb0E", r"sample_sliced_kernelILi8ELb0ELb1ELb0ELb1E",
                        r"light_kernel", r"noise_kernel", r"layout_rows_kernel"]
KEY = ["UBLKCP", "SYNCS", "IDP", "LDS", "STS", "ATOMS", "LOP3", "VOTE", "SHFL", "PRMT", "FADD", "LDG", "STG", "BAR", "BRA", "BRX", "UTMALDG", "UTCHMMA", "HMMA"]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")), key=os.path.getsize)
names = subprocess.run(["cuobjdump", "-elf", cubin], capture_output=True, text=True).stdout
funcs = sorted(set(re.findall(r"\.text\.(\S+)", names)))
print(f"SASS opcode histogram, `{os.path.relpath(SO, ROOT)}` (sm_100a, static instruction counts)\n")
for rx in WANT:
    for fn in [f for f in funcs if re.search(rx, f)]:
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, cubin], capture_output=True, text=True).stdout
        ops = collections.Counter()
        for ln in sass.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", ln)
            if m:
                ops[m.group(1)] += 1
                if m.group(1) in ("LDS", "IDP", "UBLKCP", "SYNCS"):
                    ops[m.group(1) + m.group(2)] += 0  # keep the variants visible below
        demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        total = sum(ops.values())
        print(f"### `{demangled}` -- {total} instructions ({total * 16 // 1024} KB)\n")
        print("| " + " | ".join(k for k in KEY if ops.get(k)) + " | other |")
        print("|" + "---|" * (sum(1 for k in KEY if ops.get(k)) + 1))
        print("| " + " | ".join(str(ops[k]) for k in KEY if ops.get(k)) + f" | {total - sum(ops[k] for k in KEY if ops.get(k))} |")
        top = ", ".join(f"{k} {v}" for k, v in ops.most_common(12) if v)
        print(f"\ntop opcodes: {top}\n")
