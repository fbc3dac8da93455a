#!/usr/bin/env python3
"""ncu `--page source --csv` -> the hottest SASS lines with their dominant stall reasons
(samples, share, instructions executed, SASS, top two stalls).  Usage: ncu_hot_sass.py in.csv out.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out = []
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col["# Samples"]])
    except ValueError:
        continue
    total += n
    st = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)
    out.append((n, r[col["Instructions Executed"]], r[col["Source"]].strip(),
                "%s=%d" % (st[0][1], st[0][0]), "%s=%d" % (st[1][1], st[1][0])))
out.sort(reverse=True)
with open(sys.argv[2], "w") as f:
    f.write("samples,pct,instructions_executed,sass,top_stall,second_stall\n")
    for n, ie, src, a, b in out[:60]:
        f.write("%d,%.2f,%s,%s,%s,%s\n" % (n, 100.0 * n / max(total, 1), ie, src.replace(",", ";"), a, b))
    f.write("# total samples %d\n" % total)
    agg = {}
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        for s in stalls:
            try:
                agg[s] = agg.get(s, 0) + int(r[col[s]] or 0)
            except ValueError:
                pass
    tot = sum(agg.values()) or 1
    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        f.write("# %s %.1f %%\n" % (s, 100.0 * v / tot))
