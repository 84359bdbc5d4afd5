"""Summarise an `ncu --page source --csv` dump: stall samples per source line / per reason.
usage: python scripts/ncu_stalls.py dump.csv [top]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per_reason = defaultdict(int)
lines = []
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    total += n
    rs = {s: int(r[col[s]] or 0) for s in stalls}
    for s, v in rs.items():
        per_reason[s] += v
    lines.append((n, r[col["Source"]], rs))
print("total samples", total)
for s, v in sorted(per_reason.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:28s} {v:9d} {100.0 * v / max(total, 1):5.1f}%")
print("top instructions:")
for n, src, rs in sorted(lines, key=lambda t: -t[0])[:top]:
    main = max(rs.items(), key=lambda kv: kv[1])
    print(f"  {n:8d} {100.0 * n / max(total, 1):5.1f}%  {main[0]:18s} {src[:110]}")
