"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
k = 0
for r in rows[1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    k += 1
    if k <= skip:
        continue
    name = r[ci["Kernel Name"]].split("(")[0].replace("ddpm::", "")[:100]
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{k - skip} launches, {tot:.1f} us total (ncu times are cold-cache and serialised: compare SHARES)")
for name, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  avg={v[1] / v[0]:7.1f} us  {name}")
