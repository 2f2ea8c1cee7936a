"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py launches.csv [top_n]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
    k = row["Kernel Name"].split("(")[0][-60:]
    agg[k][0] += 1; agg[k][1] += v; tot += v
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}% n={c:4d} avg={t/c:9.1f} us  {k}")
print(f"total {tot/1e3:.3f} ms")
