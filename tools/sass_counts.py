#!/usr/bin/env python
"""SASS mnemonic counts per kernel of the built library (evidence for tcgen05 / TMA / tensor-core / packed-FMA use).

    python tools/sass_counts.py > profiles/r02_sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaussian_process_odes_b200", "libgpode_b200.so")
PAT = re.compile(r'\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTCCP|UTCBAR|UBLKCP|UTMALDG|HMMA|FFMA2|FMUL2|FADD2|MUFU\.COS|MUFU\.SIN|'
                 r'MUFU\.EX2|DFMA|SYNCS|ATOM|RED|ATOMS)\b')

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        # mnemonics only (skip the encoding comment)
        code = line.split("/*", 2)[-1] if line.count("/*") >= 2 else line
        code = line[line.find("*/") + 2:] if "*/" in line else line
        for mm in PAT.finditer(code.split("/*")[0]):
            counts[cur][mm.group(1)] += 1
names = list(counts)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of libgpode_b200.so (cuobjdump -sass, all cubins sm_100a)")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA bulk copy), HMMA = mma.sync,")
print("# FFMA2 = fma.rn.f32x2, SYNCS = mbarrier ops, ATOM / RED / ATOMS = atomics (global / global reduction / shared)")
rows = []
for n, d in zip(names, dem):
    c = counts[n]
    short = re.sub(r'\(anonymous namespace\)::', '', d)
    short = re.sub(r'\(.*', '', short).replace("void ", "")
    rows.append((short, c))
for short, c in sorted(rows):
    print("%-46s %s" % (short[:46], " ".join("%s=%d" % kv for kv in sorted(c.items())) or "-"))
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print("# total: " + " ".join("%s=%d" % kv for kv in sorted(tot.items())))
