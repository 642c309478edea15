"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by (kernel, grid)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    k = (r['Kernel Name'].split('(')[0][:48], r['Grid Size'], r['Block Size'])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r['Metric Value']) / 1000
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1000:.2f} ms over {len(rows)} launches")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{a[1]:10.1f} us {100*a[1]/tot:5.1f}% {a[0]:5d} x {a[1]/a[0]:8.1f} us  {k[0]} grid={k[1]} block={k[2]}")
