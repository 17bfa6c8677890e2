"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per-kernel launches, time, share of the step, DRAM bytes.  usage: launch_summary.py launches.csv out.json"""
import csv
import json
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: defaultdict(float))
ids = defaultdict(set)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]])
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"(vs::)?(<?unnamed>?::)", "", name)
    v = float(r[ci["Metric Value"]].replace(",", ""))
    mult = {"ns": 1e-3, "us": 1, "ms": 1e3, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ci["Metric Unit"]], 1)
    agg[name][r[ci["Metric Name"]]] += v * mult
    ids[name].add(r[ci["ID"]])
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
out = {}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n, t = len(ids[k]), a["gpu__time_duration.sum"]
    out[k] = {"launches": n, "time_us": round(t, 1), "share": round(t / tot, 4),
              "dram_read_MB": round(a["dram__bytes_read.sum"] / 1e6, 1), "dram_write_MB": round(a["dram__bytes_write.sum"] / 1e6, 1),
              "dram_MB_per_launch": round((a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / 1e6 / n, 2)}
    print(f"{k[:46]:46s} n={n:4d} {t:9.1f}us {100 * t / tot:5.1f}%  rd {a['dram__bytes_read.sum'] / 1e6:9.1f} MB wr {a['dram__bytes_write.sum'] / 1e6:9.1f} MB")
print("total us", round(tot, 1), "launches", sum(len(v) for v in ids.values()))
if len(sys.argv) > 2:
    json.dump({"source": "ncu --nvtx --nvtx-include vs_timed_eager/ --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
               "dram__bytes_write.sum --clock-control none python bench.py --steps 1 --warmup 3 --no-graph (one eager step; "
               "per-launch times are serialised / cold-cache: use the SHARES)", "total_us": round(tot, 1), "kernels": out},
              open(sys.argv[2], "w"), indent=1)
