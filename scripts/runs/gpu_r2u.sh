#!/bin/bash
run() {  # tag cfg n devs port extra-env
  local tag=$1 cfg=$2 n=$3 devs=$4 port=$5; shift 5
  env CUDA_VISIBLE_DEVICES=$devs "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $cfg --steps 20 --warmup 5 --check-grads > gpurun_out/bench_r2u_${tag}.json 2> gpurun_out/bench_r2u_${tag}.err
  echo "$tag rc=$? $(python -c "
import json,sys
try:
  d=json.loads(open('gpurun_out/bench_r2u_${tag}.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), 'sus', round(d['sustained']['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('allreduce_early_buckets'), (d.get('grad_check') or {}).get('worst_rel_l2'))
except Exception as e: print('ERR', e)
")"
}
run c2_8gpu_nb1 c2 8 0,1,2,3,4,5,6,7 29601 WN_AR_BUCKETS=1
run c2_8gpu_nb2 c2 8 0,1,2,3,4,5,6,7 29602 WN_AR_BUCKETS=2
run c2_8gpu_nb3 c2 8 0,1,2,3,4,5,6,7 29603 WN_AR_BUCKETS=3
run c2_8gpu_nb4 c2 8 0,1,2,3,4,5,6,7 29604 WN_AR_BUCKETS=4
run c5_8gpu_nb3 c5 8 0,1,2,3,4,5,6,7 29605 WN_AR_BUCKETS=3
