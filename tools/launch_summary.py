"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count, total us, share."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    a = agg.setdefault(row['Kernel Name'][:int(sys.argv[2]) if len(sys.argv) > 2 else 100], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%10.1f us %5.1f%% x%-4d %s' % (t, 100 * t / tot, n, k))
print('total %.1f us' % tot)
