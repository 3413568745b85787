"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else None
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    if only and not re.search(only, name):
        continue
    agg[name][0] += 1; agg[name][1] += ms; tot += ms
print(f"total {tot:.2f} ms over {sum(n for n, _ in agg.values())} launches")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{ms:9.2f} ms {100 * ms / tot:5.1f}%  n={n:5d}  avg {1e3 * ms / n:9.1f} us  {k[:80]}")
