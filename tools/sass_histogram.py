#!/usr/bin/env python
"""Per-kernel histogram of the SASS opcodes that prove a Blackwell-native path (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, mbarrier -> SYNCS, legacy tensor path -> HMMA (must be
absent).  Reads the shipped library with cuobjdump; no GPU needed.

usage: python tools/sass_histogram.py [path/to/libim2im_uq.so] > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "im2im_uq_b200", "lib", "libim2im_uq.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "ATOMS", "REDG", "RED", "ATOMG"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"im2im::\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for k in KEYS:
            if op == k or (k in ("UTCHMMA", "UTCQMMA") and op.startswith(k)):
                kernels[cur][k] += 1
print(f"SASS opcode histogram of {os.path.relpath(lib, ROOT)} (architectures in the fatbin: {', '.join(arch)})")
print("kernels with tensor-core / TMA / TMEM instructions first; counts are static instruction counts\n")
hdr = f"{'kernel':72s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEYS)
print(hdr)
tot = collections.Counter()
rows = sorted(kernels.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["UTMALDG"] * 10 + kv[1]["UBLKCP"]))
for name, c in rows:
    print(f"{name[:72]:72s} {c['_total']:7d} " + " ".join(f"{c[k]:7d}" for k in KEYS))
    tot.update(c)
print(f"\n{'TOTAL':72s} {tot['_total']:7d} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
print("\nHMMA (mma.sync / wmma, the legacy tensor path) must be 0:", tot["HMMA"])
