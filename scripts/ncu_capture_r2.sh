#!/bin/bash
# ncu --set full captures of the round-2 kernels (one GPU): third eager step of scripts/prof_step.py.  The reports are summarised
# on the box (key counters + top SASS lines by stall samples) — gpurun brings back at most 64 MiB.
# usage (on the GPU box): scripts/ncu_capture_r2.sh <tag>
TAG=${1:-r2p}
cap() {  # cfg kernel-regex name
  ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_${TAG}_$3 python scripts/prof_step.py $1 3 > gpurun_out/ncu_${TAG}_$3.log 2>&1
  tail -n 1 gpurun_out/ncu_${TAG}_$3.log
  { echo "# ncu --set full --clock-control none, kernel regex $2, third eager step of scripts/prof_step.py $1 (scripts/ncu_capture_r2.sh)"; scripts/ncu_summary.sh gpurun_out/prof_${TAG}_$3.ncu-rep; echo; echo "# top SASS lines by warp-stall samples"; python scripts/ncu_hot.py gpurun_out/prof_${TAG}_$3.ncu-rep 0 30; } > gpurun_out/ncu_full_${TAG}_$3.txt 2>&1
  rm -f gpurun_out/prof_${TAG}_$3.ncu-rep
}
rm -f gpurun_out/*.ncu-rep
python scripts/prof_step.py c2 3 > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
cap c2 tc_stack_bwd c2_stackbwd
cap c3 tc_stack_bwd c3_stackbwd
cap c3 tc_wgrad_group_kernel c3_wgroup
