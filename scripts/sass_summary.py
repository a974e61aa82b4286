"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP; legacy HMMA must be absent) in libwavenet_b200.so.

  python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'wavenets_b200', 'libwavenet_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([
  ('UTCHMMA', r'\bUTCHMMA\b(?!\.2CTA)'), ('UTCHMMA.2CTA', r'\bUTCHMMA\.2CTA'), ('UTCBAR', r'\bUTCBAR'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'),
  ('UTMALDG', r'\bUTMALDG'), ('UTMASTG', r'\bUTMASTG'), ('UTMAPF', r'\bUTMAPF|UTMACCTL'), ('UBLKCP', r'\bUBLKCP'), ('SYNCS', r'\bSYNCS'),
  ('MUFU.TANH', r'\bMUFU\.TANH'), ('HMMA(legacy)', r'\bHMMA\b'), ('FFMA', r'\bFFMA'), ('RED/ATOM', r'\b(RED|ATOMG|ATOMS)\b')])
cur, counts, order = None, {}, []
for line in sass.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    cur = m.group(1)
    counts[cur] = collections.Counter()
    order.append(cur)
    continue
  if cur is None:
    continue
  for k, p in pats.items():
    if re.search(p, line):
      counts[cur][k] += 1
dem = subprocess.run(['cu++filt'] + order, capture_output=True, text=True).stdout.splitlines() if order else []
names = {o: (d if d else o) for o, d in zip(order, dem)} if len(dem) == len(order) else {o: o for o in order}
arch = re.findall(r'arch = (sm_\w+)', sass)
print(f'# {os.path.relpath(lib, ROOT)}: {len(order)} kernels, arch {sorted(set(arch))}; cuobjdump -sass, matches per kernel')
print('# ' + ' '.join(f'{k:>13s}' for k in pats) + '  kernel')
tot = collections.Counter()
for o in order:
  c = counts[o]
  tot.update(c)
  short = re.sub(r'\(.*', '', names[o])
  if len(short) > 110:
    short = short[:107] + '...'
  print('  ' + ' '.join(f'{c[k]:13d}' for k in pats) + '  ' + short)
print('# ' + ' '.join(f'{tot[k]:13d}' for k in pats) + '  TOTAL')
tc = [o for o in order if counts[o]['UTCHMMA'] + counts[o]['UTCHMMA.2CTA'] > 0]
print(f'# kernels issuing tcgen05.mma: {len(tc)}; with cta_group::2: {sum(1 for o in tc if counts[o]["UTCHMMA.2CTA"] > 0)}; legacy HMMA instructions: {tot["HMMA(legacy)"]}')
