"""Development aid: aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: launch_summary.py launches.csv [split-marker-kernel-substring]"""
import collections
import csv
import re
import sys


def main():
    path, marker = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = [(re.sub(r"\(.*", "", r["Kernel Name"]), float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Grid Size"], r["Block Size"])
            for r in csv.DictReader(lines)]
    cuts = [i for i, r in enumerate(rows) if marker and marker in r[0]] or [0]
    cuts.append(len(rows))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        agg = collections.defaultdict(lambda: [0, 0.0])
        for name, v, u, _, _ in rows[lo:hi]:
            agg[name[:90]][0] += 1
            agg[name[:90]][1] += v * scale[u[:2]]
        print(f"-- launches {lo}..{hi}: {sum(v[1] for v in agg.values()):.2f} ms")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {v[1]:9.3f} ms {v[0]:5d}x  {k}")


if __name__ == "__main__":
    main()
