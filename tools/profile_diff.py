"""Compare per-shape profile CSVs (bench.py --tag): python tools/profile_diff.py base.csv other.csv [...]"""
import csv
import sys


def load(p):
    d = {}
    for r in csv.DictReader(open(p)):
        key = (r["category"], r["m"], r["n"], r["k"], r["work_per_launch"])
        d[key] = (float(r["total_ms"]) / int(r["launches"]), int(r["launches"]), float(r["work_per_launch"]))
    return d


def main():
    tabs = [load(p) for p in sys.argv[1:]]
    base = tabs[0]
    keys = sorted(base, key=lambda k: -base[k][0] * base[k][1])
    tot = [0.0] * len(tabs)
    for k in keys:
        per, n, w = base[k]
        line = f"{k[0]} {k[1]:>7} {k[2]:>6} {k[3]:>6} n={n:4d} {per * 1e3:8.1f}us {w / per / 1e9:7.1f}"
        for i, t in enumerate(tabs):
            if k in t:
                tot[i] += t[k][0] * t[k][1]
            if i and k in t:
                line += f" | {t[k][0] * 1e3:8.1f}us {100 * (t[k][0] / per - 1):+6.1f}%"
        if per * n > 0.02 * sum(v[0] * v[1] for v in base.values()) / 10:
            print(line)
    print("totals(ms over all steps):", [round(x, 2) for x in tot])


main()
