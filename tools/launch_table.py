"""Per-(kernel, grid) table of an ncu launch list (gpu__time_duration.sum CSV)."""
import collections, csv, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(io.StringIO(''.join(lines))))
agg = collections.OrderedDict()
for r in rows:
    name = r['Kernel Name'].split('(')[0][-48:]
    grid = r.get('Grid Size', '')
    try:
        v = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    if r.get('Metric Unit', 'ns') in ('us', 'usecond'):
        v *= 1e3
    a = agg.setdefault((name, grid), [0, 0.0]); a[0] += 1; a[1] += v
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for (name, grid), (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{name:48s} {grid:>18s} n={c:3d} avg={t/c/1e3:9.1f} us total={t/1e3:9.1f}")
