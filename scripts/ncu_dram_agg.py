"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log per kernel.
usage: python scripts/ncu_dram_agg.py gpurun_out/f2_dram.csv [out.csv]"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
c = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: defaultdict(float))
ids = defaultdict(set)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[c["Kernel Name"]]).replace("void ", "").replace("ic::", "").replace("(anonymous namespace)::", "")
    val = float(r[c["Metric Value"]].replace(",", ""))
    unit = r[c["Metric Unit"]]
    m = r[c["Metric Name"]]
    if m == "gpu__time_duration.sum":
        val *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    else:
        val *= {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(unit, 1e-9)
    agg[name][m] += val
    ids[name].add(r[c["ID"]])
out = [("kernel", "launches", "ms", "dram_read_GB", "dram_write_GB")]
tot = [0.0, 0.0, 0.0]
for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    v = (d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"])
    tot = [a + b for a, b in zip(tot, v)]
    out.append((name, len(ids[name]), f"{v[0]:.2f}", f"{v[1]:.2f}", f"{v[2]:.2f}"))
out.append(("total", "", f"{tot[0]:.2f}", f"{tot[1]:.2f}", f"{tot[2]:.2f}"))
w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
w.writerows(out)
