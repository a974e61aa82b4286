#!/bin/bash
# ncu evidence for profiles/ (one GPU): launch list of one C2 step + --set full captures of the GEMM-class kernels.
# usage (on the GPU box): scripts/ncu_capture.sh <tag>
TAG=${1:-r1e}
python scripts/prof_step.py c2 3 > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
PER_STEP=$(grep -o "launches/step [0-9]*" gpurun_out/plain_$TAG.log | grep -o "[0-9]*$")
# launch list (recipe: --metrics gpu__time_duration.sum --clock-control none); setup = 2 launches (one-launch re-pack), then PER_STEP per step
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2 + 2 * PER_STEP)) -c $PER_STEP --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_list.log 2>&1
# one launch of each GEMM-class kernel, second step
ncu --set full --clock-control none --import-source on -k regex:tc_block_fwd --launch-skip 40 --launch-count 1 -f -o gpurun_out/prof_${TAG}_blockfwd python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_a.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:tc_conv_gemm_staged|tc_wgrad" --launch-skip 120 --launch-count 6 -f -o gpurun_out/prof_${TAG}_bwd python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_b.log 2>&1
tail -n 2 gpurun_out/ncu_${TAG}_a.log; tail -n 2 gpurun_out/ncu_${TAG}_b.log; echo per_step $PER_STEP
