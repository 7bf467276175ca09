#!/usr/bin/env python
"""Print an ncu launch list (csv from --metrics ... --csv --log-file) as one line per kernel launch."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
d = defaultdict(dict); names = {}
for r in rows[1:]:
    if len(r) < len(hdr): continue
    k = int(r[idx["ID"]]); names[k] = r[idx["Kernel Name"]][:34]
    d[k][r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
tot = 0; rd = 0; wr = 0
for k in sorted(d):
    m = d[k]; t = m["gpu__time_duration.sum"] / 1e6; tot += t
    rd += m.get("dram__bytes_read.sum", 0); wr += m.get("dram__bytes_write.sum", 0)
    print(f"{k:3d} {names[k]:34s} grid {int(m.get('launch__grid_size',0)):5d} t {t:7.3f} ms rd {m.get('dram__bytes_read.sum',0)/1e9:6.2f} GB "
          f"wr {m.get('dram__bytes_write.sum',0)/1e9:6.2f} GB l2hit {m.get('lts__t_sector_hit_rate.pct',0):5.1f} "
          f"inst {m.get('sm__inst_executed.sum',0)/1e9:6.2f}G issue {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f}%")
print(f"total {tot:.3f} ms, dram read {rd/1e9:.1f} GB write {wr/1e9:.1f} GB")
