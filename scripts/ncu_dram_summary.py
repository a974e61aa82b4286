"""ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv`) of ONE
training step -> profiles/dram_<config>.json: per-kernel launches, DRAM bytes and device time of the step.  bench.py reads that
file for `roofline.traffic` / `roofline.traffic_by_kernel` (bytes per step per kernel against the algorithmic bytes).

  python scripts/ncu_dram_summary.py gpurun_out/launches_r2_c2.csv <per_step> c2 bf16 profiles/dram_c2.json "<how it was captured>"
"""
import collections, csv, json, re, sys

path, per_step, config, precision, out = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
source = sys.argv[6] if len(sys.argv) > 6 else path
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
rows = list(launch.values())[-per_step:] if per_step > 0 else list(launch.values())
tot = collections.OrderedDict()
for d in rows:
  short = re.sub(r'\(.*', '', d['name']).replace('void ', '')
  t = tot.setdefault(short, {'launches': 0, 'dram_read': 0.0, 'dram_write': 0.0, 'ms': 0.0})
  t['launches'] += 1; t['dram_read'] += d['rd']; t['dram_write'] += d['wr']; t['ms'] += d['us'] / 1000.0
res = {'config': config, 'precision': precision, 'source': source, 'launches_per_step': len(rows),
       'ncu_ms_per_step_serialised_cold_cache': sum(t['ms'] for t in tot.values()),
       'dram_bytes_per_step': sum(t['dram_read'] + t['dram_write'] for t in tot.values()),
       'kernels': dict(sorted(tot.items(), key=lambda kv: -kv[1]['ms']))}
with open(out, 'w') as f:
  json.dump(res, f, indent=1)
print(f'{out}: {len(rows)} launches, {res["ncu_ms_per_step_serialised_cold_cache"]:.3f} ms, {res["dram_bytes_per_step"] / 1e9:.2f} GB')
