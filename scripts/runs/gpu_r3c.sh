#!/bin/bash
# final validation of the re-entry session: full GPU suite, default bench (C2), C4 (conditioning MLP + mixture loss), smoke
set -x
( time python -m pytest tests -m gpu -q --maxfail=6 ) > gpurun_out/r3c_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r3c_tests.log
python bench.py > gpurun_out/bench_r3c_c2.json 2> gpurun_out/bench_r3c_c2.err; echo "bench default rc=$?"
python bench.py --config c4 > gpurun_out/bench_r3c_c4.json 2> gpurun_out/bench_r3c_c4.err; echo "bench c4 rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -4
python - <<'P'
import json
for c in ('c2', 'c4'):
  d = json.loads(open(f'gpurun_out/bench_r3c_{c}.json').read().strip().splitlines()[-1])
  k = d['kernels']
  print(c, round(d['value']), d['ms_per_step'], d['sustained']['ms_per_step'], d['gpu_launches'], 'cond_fwd', k.get('cond_fwd'), 'cond_bwd', k.get('cond_bwd'), 'loss', k.get('loss'))
P
