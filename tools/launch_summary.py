"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised: compare shares)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except Exception:
        continue
    unit = r[ix["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
    k = r[ix["Kernel Name"]][:64]
    agg[k][0] += 1; agg[k][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, {tot:.2f} ms total")
for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:10]:
    print(f"  {k:66s} n={n:4d}  {ms:9.3f} ms  {100*ms/tot:5.1f}%")
