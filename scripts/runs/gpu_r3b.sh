#!/bin/bash
# loss kernels after the rewrite (mixture loss: G lanes per row; softmax CE: cheap clip detection): parity, then timing
set -x
( time python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity_fp32.py tests/test_gpu_parity_bf16.py tests/test_gpu_sampling.py tests/test_gpu_configs.py -m gpu -q --maxfail=6 ) > gpurun_out/r3b_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r3b_tests.log
python bench.py --config c4 > gpurun_out/bench_r3b_c4.json 2> gpurun_out/bench_r3b_c4.err; echo "bench c4 rc=$?"
python bench.py > gpurun_out/bench_r3b_c2.json 2> gpurun_out/bench_r3b_c2.err; echo "bench c2 rc=$?"
python - <<'P'
import json
for c in ('c4', 'c2'):
  d = json.loads(open(f'gpurun_out/bench_r3b_{c}.json').read().strip().splitlines()[-1])
  print(c, round(d['value']), d['ms_per_step'], d['roofline_hbm']['loss'])
P
