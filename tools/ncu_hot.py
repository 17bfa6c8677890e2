"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: ncu_hot.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions", len(body))
agg = {s: sum(int(r[ci[s]] or 0) for r in body) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(body, key=lambda r: -int(r[ci["# Samples"]] or 0))[:n]
for r in sorted(top, key=lambda r: int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else 0):
    s = int(r[ci["# Samples"]] or 0)
    why = sorted(((int(r[ci[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{r[ci['Address']][-5:]} {100.0 * s / tot:5.1f}% ex={r[ci['Instructions Executed']]:>8} {r[ci['Source']][:70]:70s} {why}")
