"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("xn::", "")
    agg[n][0] += 1
    agg[n][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, total {tot/1e6:.3f} ms (unit {rows[0]['Metric Unit']}; cold-cache, serialised: compare shares)")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e6:9.3f} ms {v[0]:5d}x {100*v[1]/tot:5.1f}%  {n[:100]}")
