#!/bin/bash
run() {  # tag cfg n devs port extra-env
  local tag=$1 cfg=$2 n=$3 devs=$4 port=$5; shift 5
  if [ $n -eq 1 ]; then
    env CUDA_VISIBLE_DEVICES=$devs "$@" python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2zd_${tag}.json 2> gpurun_out/bench_r2zd_${tag}.err
  else
    env CUDA_VISIBLE_DEVICES=$devs "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $cfg --steps 20 --warmup 5 --check-grads > gpurun_out/bench_r2zd_${tag}.json 2> gpurun_out/bench_r2zd_${tag}.err
  fi
  echo "$tag rc=$? $(python -c "
import json,sys
try:
  d=json.loads(open('gpurun_out/bench_r2zd_${tag}.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), 'sus', round(d['sustained']['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('allreduce_early_buckets'), (d.get('grad_check') or {}).get('worst_rel_l2'))
except Exception as e: print('ERR', e)
")"
}
run c2_8gpu c2 8 0,1,2,3,4,5,6,7 29601
run c3_8gpu c3 8 0,1,2,3,4,5,6,7 29602
run c4_8gpu c4 8 0,1,2,3,4,5,6,7 29603
run c5_8gpu c5 8 0,1,2,3,4,5,6,7 29604
( CUDA_VISIBLE_DEVICES=6,7 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2zd_multi.log 2>&1; echo "multi rc=$?"; tail -2 gpurun_out/r2zd_multi.log ) &
run c2_1gpu c2 1 0 0 & run c3_1gpu c3 1 1 0 & run c4_1gpu c4 1 2 0 & run c5_1gpu c5 1 3 0 & wait
