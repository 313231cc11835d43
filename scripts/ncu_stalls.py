#!/usr/bin/env python
"""Top stall locations (SASS) of the first kernel in an .ncu-rep.  Usage: python scripts/ncu_stalls.py rep [ntop]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) > iSm and r[iSm].isdigit()]
tot = sum(int(r[iSm]) for r in body)
print('total samples', tot)
agg = {}
for r in body:
    for i in stall:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {k:28s} {v:6d} {100 * v / tot:5.1f}%")
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for r in sorted(body, key=lambda r: -int(r[iSm]))[:ntop]:
    why = sorted([(int(r[i] or 0), hdr[i]) for i in stall], reverse=True)[:2]
    print(r[0][-5:], f"{r[iS][:64]:64s}", f"{r[iSm]:>5s} {r[iE]:>8s}", why)
