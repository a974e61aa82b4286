#!/bin/bash
# ncu evidence for profiles/ (one GPU), tree with the grouped weight-gradient launch: launch list of one C2 step +
# --set full captures of the grouped wgrad kernel (final launch), its finish, and one gate adjoint / dgrad.
# usage (on the GPU box): scripts/ncu_capture2.sh <tag>
TAG=${1:-r1g}
python scripts/prof_step.py c2 3 > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
PER_STEP=$(grep -o "launches/step [0-9]*" gpurun_out/plain_$TAG.log | grep -o "[0-9]*$")
# setup = 2 launches (one-launch re-pack), then PER_STEP per step; third step = side launches active (plan exists, eager)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip $((2 + 2 * PER_STEP)) -c $PER_STEP --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_list.log 2>&1
WN_SIDE_STREAM=0 ncu --set full --clock-control none --import-source on -k "regex:tc_wgrad_group" --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_${TAG}_wgroup python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_a.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:tc_stack_fwd" --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_${TAG}_stackfwd python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_b.log 2>&1
tail -n 2 gpurun_out/ncu_${TAG}_a.log; tail -n 2 gpurun_out/ncu_${TAG}_b.log; echo per_step $PER_STEP
