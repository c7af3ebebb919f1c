#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: loops (runs of instructions with the same executed
count), their share of the stall samples, cycles of sub-partition time per trip, and the stall mix; optionally the
instructions of one loop.   tools/sass_hot.py file.csv [loop_index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
data, seen = [], set()
for r in rows[h + 1:]:   # first captured launch only (an export may hold several launches of the kernel)
    if len(r) < len(hdr):
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    data.append(r)
ix = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]


def I(r, k):
    try:
        return int(r[ix[k]])
    except ValueError:
        return 0


groups = []
for i, r in enumerate(data):
    e = I(r, "Instructions Executed")
    if not groups or groups[-1]["e"] != e:
        groups.append(dict(e=e, rows=[]))
    groups[-1]["rows"].append(r)
tot = sum(I(r, "# Samples") for r in data)
print("samples", tot, "warp instructions", sum(I(r, "Instructions Executed") for r in data))
big = [g for g in groups if sum(I(r, "# Samples") for r in g["rows"]) > 0.01 * tot]
for gi, g in enumerate(big):
    s = sum(I(r, "# Samples") for r in g["rows"])
    mix = {k[6:]: sum(I(r, k) for r in g["rows"]) for k in stalls}
    top = sorted(mix.items(), key=lambda kv: -kv[1])[:5]
    nm = sum("MUFU" in r[1] for r in g["rows"])
    print("loop %d: exec %d, %d instr (%d MUFU), share %.3f, stalls %s" % (
        gi, g["e"], len(g["rows"]), nm, s / tot, " ".join("%s=%.2f" % (k, v / s) for k, v in top)))
if len(sys.argv) > 2:
    for r in big[int(sys.argv[2])]["rows"]:
        st = {k[6:]: I(r, k) for k in stalls if I(r, k) > 0.004 * tot / 10}
        print("%6d  %-70s %s" % (I(r, "# Samples"), r[1][:70].strip(), st))
