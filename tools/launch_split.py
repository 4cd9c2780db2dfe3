#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list, starting at the
last launch whose name contains --from (default: k_flatten = the last BVH build in the run)."""
import argparse
import collections
import csv

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--from", dest="start", default="k_flatten")
args = ap.parse_args()
hdr, seq = None, []
for r in csv.reader(open(args.csv)):
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
    seq.append((d["Kernel Name"].split("(")[0][-48:], v))
idx = [i for i, (k, _) in enumerate(seq) if args.start in k]
seq = seq[idx[-1]:] if idx else seq
agg = collections.OrderedDict()
for k, v in seq:
    a = agg.setdefault(k, [0.0, 0])
    a[0] += v
    a[1] += 1
tot = sum(a[0] for a in agg.values())
for k, (v, c) in agg.items():
    print(f"{k:50s} {c:4d} launches {v:9.1f} us {100 * v / tot:5.1f} %")
print(f"{'total':50s} {len(seq):4d} launches {tot:9.1f} us")
