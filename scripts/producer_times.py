"""Per-step durations of the producer sweeps from an ncu launch list (debug aid): python scripts/producer_times.py file.csv [every]"""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
every = int(sys.argv[2]) if len(sys.argv) > 2 else 3
out, step = [], []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('rtp::', '')
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1000 if row['Metric Unit'] == 'ns' else v
    if name.startswith('fluidPredict'):
        if step:
            out.append(step)
        step = []
    if name.startswith('densityLambda') or name.startswith('vorticity'):
        step.append(round(v))
out.append(step)
for i, s in enumerate(out):
    if i % every == 0:
        print(i, s)
