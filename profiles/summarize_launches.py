#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names = rows[hdr]
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(names):
            continue
        d = dict(zip(names, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        key = (d["Kernel Name"].split("(")[0], d["Block Size"], d["Grid Size"])
        ns = float(d["Metric Value"].replace(",", ""))
        if d["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        cnt, tot = agg.get(key, (0, 0.0))
        agg[key] = (cnt + 1, tot + ns)
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {total / 1e6:.3f} ms of device time "
          "(ncu per-launch times are cold-cache and serialised: compare shares)")
    print(f"{'launches':>8} {'total_ms':>10} {'avg_us':>10} {'share':>7}  kernel  block grid")
    for key, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{cnt:8d} {tot / 1e6:10.3f} {tot / cnt / 1e3:10.1f} {tot / total * 100:6.1f}%  "
              f"{key[0]}  {key[1]} {key[2]}")


if __name__ == "__main__":
    main(sys.argv[1])
