"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one recurrent iteration + totals."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') == 'gpu__time_duration.sum':
        rows.append((row['Kernel Name'].split('(')[0].replace('<unnamed>::', ''), float(row['Metric Value'].replace(',', '')), row.get('Grid Size')))
names = [r[0] for r in rows]
fi = [i for i, n in enumerate(names) if 'flow_init' in n]
if len(fi) > 2:
    print("one recurrent iteration:")
    for k in range(fi[1], fi[2]):
        print(f"  {rows[k][0][-40:]:42s} {rows[k][1] / 1e3:9.1f} us  grid {rows[k][2]}")
    print(f"  iteration total {sum(r[1] for r in rows[fi[1]:fi[2]]) / 1e3:.1f} us")
agg = collections.OrderedDict()
for k, v, g in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("totals over the captured launches:")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
    print(f"  {k[-45:]:45s} n={n:4d} total={t / 1e3:9.1f} us avg={t / n / 1e3:8.1f} share={t / tot * 100:5.1f}%")
