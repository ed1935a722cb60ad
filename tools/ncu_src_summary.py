#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall samples per mbarrier wait loop and per opcode class.
   usage: ncu -i rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_src_summary.py src.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; cols = {h: i for i, h in enumerate(hdr)}
seen, ins = set(), []
for r in rows[hi + 1:]:
    if not r or r[0] in seen or not r[0].startswith("0x"):
        continue
    seen.add(r[0])
    ins.append((r[0], r[1].strip(), float(r[cols["# Samples"]] or 0), int(r[cols["Instructions Executed"]] or 0)))
tot = sum(v for _, _, v, _ in ins)
print("instructions", len(ins), "samples", tot)
for i, (a, sx, v, n) in enumerate(ins):
    if "TRYWAIT" in sx:
        s = sum(x[2] for x in ins[i:i + 3])
        if s / tot > 0.004:
            print(f"{s:7.0f} {100 * s / tot:5.1f}%  {sx[:80]}  execs={n}")
agg = collections.Counter()
for a, sx, v, n in ins:
    t = sx.split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    agg[op.split(".")[0]] += v
print("by opcode:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in agg.most_common(16)))
