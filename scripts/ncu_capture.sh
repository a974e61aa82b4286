#!/bin/bash
# ncu evidence for profiles/ (one GPU): launch list of one C2 step + --set full captures of the GEMM-class kernels.
# usage (on the GPU box): scripts/ncu_capture.sh <tag>
TAG=${1:-r1c}
python scripts/prof_step.py c2 3 > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
# launch list (recipe: --metrics gpu__time_duration.sum --clock-control none); setup = ~560 weight-pack launches, then 273 per step
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1106 -c 273 --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_list.log 2>&1
K='regex:tc_conv_gemm_staged|tc_wgrad'
# matching launches per step: 63 forward (gate, conv1 x30 + skip/head), then head backward, then per block: wgrad 1x1, finish, gate_bwd, wgrad dilated, finish, dgrad
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 270 --launch-count 2 -f -o gpurun_out/prof_${TAG}_fwd python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 400 --launch-count 6 -f -o gpurun_out/prof_${TAG}_bwd python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_bwd.log 2>&1
tail -n 2 gpurun_out/ncu_${TAG}_fwd.log; tail -n 2 gpurun_out/ncu_${TAG}_bwd.log
