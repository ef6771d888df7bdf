"""Mean gpu__time_duration per kernel from an ncu --csv launch list (scripts/kernel_times.py file.csv)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'^void ', '', re.sub(r'\(.*', '', row['Kernel Name'])).replace('rtp::', '')
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1000 if row['Metric Unit'] == 'ns' else (v * 1000 if row['Metric Unit'] == 'ms' else v)
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for k, v in agg.items() if not k.startswith('at::'))
steps = max(len(v) for v in agg.values()) if agg else 1
for k, v in agg.items():
    if k.startswith('at::'):
        continue
    print("%-42s n=%3d mean=%8.2f us  min=%8.2f max=%8.2f share=%5.3f" % (k[:42], len(v), sum(v) / len(v), min(v), max(v), sum(v) / tot))
steps = max([len(v) for k, v in agg.items() if k.startswith(('fluidPredictKernel', 'cloudsThermoPredictKernel', 'boidsCellIdsKernel'))] or [1])
print("sum of kernel time per step ~ %.1f us (%d steps)" % (tot / steps, steps))
