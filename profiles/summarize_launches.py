"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total and average device time per kernel.

    python profiles/summarize_launches.py gpurun_out/launches.csv "comment line" [last N launches] > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import sys


def main(path, comment="", last=0):
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    rows = list(csv.reader(open(path)))
    if last:      # only the last `last` launches (steady state of a run whose first steps tune the recorded graph)
        data = [r for r in rows if len(r) >= 6 and r[0] != "ID" and r[0].isdigit()]
        keep = set(r[0] for r in data[-last:])
        rows = [r for r in rows if len(r) < 6 or r[0] == "ID" or r[0] in keep]
    for r in rows:
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = d["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg[d["Kernel Name"][:120]][0] += 1
        agg[d["Kernel Name"][:120]][1] += v
    tot = sum(v[1] for v in agg.values())
    if comment:
        print("# " + comment)
    print("# launches     total us    share    avg us  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[0]:10d} {v[1]:12.1f} {100 * v[1] / tot:7.2f}% {v[1] / v[0]:9.2f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 0)
