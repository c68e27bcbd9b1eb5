"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
tot = collections.Counter(); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print('%-44s %6s %12s %12s %7s' % ('kernel', 'n', 'total ms', 'avg us', 'share'))
for k, v in tot.most_common(40):
    print('%-44s %6d %12.3f %12.1f %6.1f%%' % (k[:44], cnt[k], v / 1e6, v / cnt[k] / 1e3, 100 * v / T))
print('total ms %.3f over %d launches' % (T / 1e6, sum(cnt.values())))
