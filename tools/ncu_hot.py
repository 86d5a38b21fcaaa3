#!/usr/bin/env python
"""Top stall-sample SASS lines of an `ncu --page source --csv --print-source sass` export.
Usage: ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = None
blocks = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; blocks.append([kern, None, []]); continue
    if r and r[0] == "Address":
        blocks[-1][1] = r; continue
    if blocks and blocks[-1][1]:
        blocks[-1][2].append(r)
for kern, hdr, body in blocks[:1]:
    ix = {h: i for i, h in enumerate(hdr)}
    si = ix["Warp Stall Sampling (All Samples)"]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    tot = sum(int(r[si] or 0) for r in body)
    print(kern, "total samples", tot, "sass lines", len(body))
    agg = {}
    for r in body:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
    order = sorted(range(len(body)), key=lambda k: -int(body[k][si] or 0))[:n]
    for k in sorted(order):
        r = body[k]
        top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{k:5d} {int(r[si] or 0):6d} {100*int(r[si] or 0)/max(tot,1):5.1f}%  {r[ix['Source']][:70]:70s} {top}")
