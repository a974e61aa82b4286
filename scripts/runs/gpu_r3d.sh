#!/bin/bash
# final tree of the re-entry session: full GPU suite, smoke, default bench, C1, ncu launch lists (time + DRAM bytes) of C2 and C4
set -x
( time python -m pytest tests -m gpu -q --maxfail=6 ) > gpurun_out/r3d_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r3d_tests.log
python __graft_entry__.py smoke 2>&1 | tail -4
python bench.py > gpurun_out/bench_r3d_c2.json 2> gpurun_out/bench_r3d_c2.err; echo "bench default rc=$?"
python bench.py --config c1 --steps 50 --warmup 5 > gpurun_out/bench_r3d_c1.json 2> gpurun_out/bench_r3d_c1.err; echo "bench c1 rc=$?"
scripts/ncu_list.sh c2 r3d
scripts/ncu_list.sh c4 r3d
python - <<'P'
import json
for c in ('c2', 'c1'):
  d = json.loads(open(f'gpurun_out/bench_r3d_{c}.json').read().strip().splitlines()[-1])
  k = d['kernels']
  print(c, round(d['value']), d['ms_per_step'], d['sustained']['ms_per_step'], d['gpu_launches'], {n: round(v['ms_per_step'], 4) for n, v in k.items() if n in ('cond_fwd', 'cond_bwd', 'loss', 'input_conv_fwd')})
P
