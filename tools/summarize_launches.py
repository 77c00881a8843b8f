import csv, collections, sys
path = sys.argv[1]
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5  # fraction of launches to skip (warm-up reps)
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
r = list(csv.DictReader(lines))
start = int(len(r) * skip)
agg = collections.OrderedDict(); tot = 0
for row in r[start:]:
    name = row['Kernel Name'].split('(')[0][-60:]
    v = float(row['Metric Value'].replace(',', '')) / 1000
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:5d} {t:10.1f} us {100*t/tot:5.1f}%  avg {t/n:8.1f} us  {k}")
print("total us", round(tot, 1))
