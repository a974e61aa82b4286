"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for
the LAST full training step found in the file (launch ids of one step = `--per-step` launches)."""
import csv, sys, re, collections
path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
rows = []
with open(path) as f:
  lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
  if r.get('Metric Name') != 'gpu__time_duration.sum':
    continue
  v = float(r['Metric Value'].replace(',', ''))
  unit = r['Metric Unit']
  us = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
  rows.append((int(r['ID']), r['Kernel Name'], us))
if per_step:
  rows = rows[-per_step:]
tot = collections.defaultdict(lambda: [0, 0.0])
for _, name, us in rows:
  short = re.sub(r'\(.*', '', name)
  short = re.sub(r'<(.*)>', lambda m: '<' + m.group(1)[:70] + '>', short)
  tot[short][0] += 1
  tot[short][1] += us
total = sum(v[1] for v in tot.values())
print(f'{len(rows)} launches, {total/1000:.3f} ms total (cold-cache, serialised)')
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
  print(f'{us/1000:9.3f} ms {100*us/total:5.1f}% {n:5d}x {us/n:8.1f} us  {name}')
