#!/bin/bash
# last 1-GPU check of the session (delta to r3d: unrolled input-conv row loop, L2 prefetch hints in the fused conditioning MLP)
set -x
( time python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity_fp32.py tests/test_gpu_configs.py tests/test_gpu_generate.py -m gpu -q --maxfail=6 ) > gpurun_out/r3e_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3e_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_r3e_c2.json 2> gpurun_out/bench_r3e_c2.err; echo "bench default rc=$?"
scripts/ncu_list.sh c2 r3e > gpurun_out/r3e_list.log 2>&1; grep -E "mapping|input_conv_fwd|softmax|launches," gpurun_out/launches_r3e_c2_summary.txt
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_r3e_c2.json').read().strip().splitlines()[-1])
print('c2', round(d['value']), d['ms_per_step'], d['sustained']['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['e2e']['value'])
P
