"""Top SASS lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (first kernel or --kernel idx)."""
import csv, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
# split per kernel
blocks, cur = [], None
for ln in out:
  if ln.startswith('"Kernel Name"'):
    cur = {'name': ln, 'lines': []}
    blocks.append(cur)
  elif cur is not None:
    cur['lines'].append(ln)
b = blocks[which]
print(b['name'][:160])
rd = list(csv.reader(b['lines']))
hdr = rd[0]
idx = {h: i for i, h in enumerate(hdr)}
rows = rd[1:]
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[idx['# Samples']] or 0) for r in rows)
print('total samples', tot)
agg = {c: sum(int(r[idx[c]] or 0) for r in rows) for c in stall_cols}
print('by reason:', ', '.join(f'{c[6:]}={v}' for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
rows_s = sorted(enumerate(rows), key=lambda ir: -int(ir[1][idx['# Samples']] or 0))[:topn]
for i, r in sorted(rows_s):
  n = int(r[idx['# Samples']] or 0)
  top = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
  print(f'{i:5d} {n:7d} {100*n/tot:5.1f}%  {r[idx["Source"]].strip()[:90]:90s} {top}')
