"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, mean us, share.
usage: python tools/ncu_launches.py launches.csv"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r["Metric Unit"]]
    agg[name].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':72s} {'n':>5s} {'mean us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:72]:72s} {len(v):5d} {sum(v) / len(v):9.1f} {100 * sum(v) / tot:5.1f}%")
print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
