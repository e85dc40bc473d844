"""Aggregate an ncu gpu__time_duration launch list: python tools/launch_agg.py file.csv [skip_regex]"""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
seq = []
for row in csv.DictReader(lines):
    name = row['Kernel Name']; v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    n = re.sub(r'void |fsb::|\(anonymous namespace\)::|<unnamed>::', '', name.split('(')[0])[:48]
    if skip and skip.search(n):
        continue
    seq.append((n, v, row.get('Grid Size')))
tot = sum(v for _, v, _ in seq)
print(f"total {tot/1e3:.2f} ms, {len(seq)} launches")
agg = collections.OrderedDict()
for n, v, g in seq:
    a = agg.setdefault((n, g), [0, 0.0]); a[0] += 1; a[1] += v
for (n, g), (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"  {v/1e3:8.3f} ms {c:4d}x {v/c:9.1f} us  {n} grid={g}")
