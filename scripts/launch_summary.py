"""Summarise an ncu `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel totals over the LAST `per_step` launches (default: whole file)."""
import csv, sys, re, collections
path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
with open(path) as f:
  lines = [l for l in f if l.startswith('"')]
launch = collections.OrderedDict()
for r in csv.DictReader(lines):
  v = float(r['Metric Value'].replace(',', ''))
  unit, name = r['Metric Unit'], r['Metric Name']
  d = launch.setdefault(int(r['ID']), {'name': r['Kernel Name'], 'us': 0.0, 'rd': 0.0, 'wr': 0.0})
  if name == 'gpu__time_duration.sum':
    d['us'] = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
  elif name.startswith('dram__bytes'):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    d['rd' if 'read' in name else 'wr'] = v * mult
rows = list(launch.values())
if per_step:
  rows = rows[-per_step:]
tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in rows:
  short = re.sub(r'\(.*', '', d['name'])
  short = re.sub(r'<(.*)>', lambda m: '<' + m.group(1)[:60] + '>', short)
  t = tot[short]
  t[0] += 1; t[1] += d['us']; t[2] += d['rd']; t[3] += d['wr']
total = sum(v[1] for v in tot.values())
print(f'{len(rows)} launches, {total/1000:.3f} ms total (cold-cache, serialised), dram read {sum(v[2] for v in tot.values())/1e9:.2f} GB write {sum(v[3] for v in tot.values())/1e9:.2f} GB')
for name, (n, us, rd, wr) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
  print(f'{us/1000:9.3f} ms {100*us/total:5.1f}% {n:5d}x {us/n:8.1f} us  rd {rd/n/1e6:7.1f} MB wr {wr/n/1e6:7.1f} MB  {name}')
