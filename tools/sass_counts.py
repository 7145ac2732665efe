#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass), written as a markdown table:
the evidence that the tensor-core kernels are tcgen05 (UTCHMMA / UTCQMMA ...), read accumulators from tensor memory
(LDTM / STTM), use bulk async copies (UBLKCP) / tensor maps (UTMALDG), and which pipes the GP kernels use (DFMA, DMMA).
usage: python tools/sass_counts.py [lib.so] > profiles/rNN_sass_counts.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "metabo_b200", "libmetabo_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "DFMA", "DMMA",
         "DADD", "DMUL", "FFMA", "HMMA", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "REDUX", "SHFL", "BAR", "ATOM", "RED",
         "MULTIMEM", "LD.E", "ST.E"]
out = subprocess.run(["cuobjdump", "-sass", lib], check=True, stdout=subprocess.PIPE).stdout.decode()
demangle = {}
kern = None
counts = collections.OrderedDict()
sizes = {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        sizes[kern] = 0
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(2)
        sizes[kern] += 1
        base = op.split(".")[0]
        counts[kern][base] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts).encode(), stdout=subprocess.PIPE).stdout.decode().splitlines()
cols = [w for w in WATCH if any(c[w] for c in counts.values())]
print("# SASS mnemonic counts per kernel (`cuobjdump -sass %s`, sm_100a)\n" % os.path.relpath(lib, ROOT))
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
tot = collections.Counter()
for (k, c), nm in zip(counts.items(), names):
    nm = re.sub(r"\(.*\)$", "", nm)
    nm = nm.replace("void ", "")
    print("| `%s` | %d | " % (nm[:90], sizes[k]) + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
    tot.update(c)
print("| **total** | %d | " % sum(sizes.values()) + " | ".join(str(tot[w]) for w in cols) + " |")
