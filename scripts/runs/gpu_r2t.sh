#!/bin/bash
# 8-GPU box: scaling runs of C2 / C4 / C5 at 1, 2, 4, 8 GPUs (runs on disjoint GPU sets go side by side)
run() {  # tag cfg n devs port extra-env
  local tag=$1 cfg=$2 n=$3 devs=$4 port=$5; shift 5
  if [ $n -eq 1 ]; then
    env CUDA_VISIBLE_DEVICES=$devs "$@" python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2t_${tag}.json 2> gpurun_out/bench_r2t_${tag}.err
  else
    env CUDA_VISIBLE_DEVICES=$devs "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $cfg --steps 20 --warmup 5 --check-grads > gpurun_out/bench_r2t_${tag}.json 2> gpurun_out/bench_r2t_${tag}.err
  fi
  echo "$tag rc=$? $(python -c "
import json,sys
try:
  d=json.loads(open('gpurun_out/bench_r2t_${tag}.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), 'sus', round(d['sustained']['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('allreduce_early_buckets'), (d.get('grad_check') or {}).get('worst_rel_l2'))
except Exception as e: print('ERR', e)
")"
}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run c2_8gpu c2 8 0,1,2,3,4,5,6,7 29601
run c2_8gpu_buckets1 c2 8 0,1,2,3,4,5,6,7 29602 WN_AR_BUCKETS=1
run c4_8gpu c4 8 0,1,2,3,4,5,6,7 29603
run c5_8gpu c5 8 0,1,2,3,4,5,6,7 29604
run c2_4gpu c2 4 0,1,2,3 29605 & run c4_4gpu c4 4 4,5,6,7 29606 & wait
run c5_4gpu c5 4 0,1,2,3 29607 & run c2_2gpu c2 2 4,5 29608 & run c4_2gpu c4 2 6,7 29609 & wait
run c5_2gpu c5 2 0,1 29610 & run c2_1gpu c2 1 2 0 & run c4_1gpu c4 1 3 0 & run c5_1gpu c5 1 4 0 & wait
