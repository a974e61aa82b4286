#!/bin/bash
# round-2 re-entry validation: full GPU suite, ncu --set full of the HBM-bound loss / input-conv kernels, bench lines, smoke
set -x
( time python -m pytest tests -m gpu -q --maxfail=6 ) > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r3a_tests.log
python bench.py > gpurun_out/bench_r3a_c2.json 2> gpurun_out/bench_r3a_c2.err; echo "bench default rc=$?"
python bench.py --config c1 --steps 50 --warmup 5 > gpurun_out/bench_r3a_c1.json 2> gpurun_out/bench_r3a_c1.err; echo "bench c1 rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -4
TAG=r3a
cap() {  # cfg kernel-regex name
  ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_${TAG}_$3 python scripts/prof_step.py $1 3 > gpurun_out/ncu_${TAG}_$3.log 2>&1
  tail -n 1 gpurun_out/ncu_${TAG}_$3.log
  { echo "# ncu --set full --clock-control none, kernel regex $2, third eager step of scripts/prof_step.py $1 (scripts/gpu_r3a.sh)"; scripts/ncu_summary.sh gpurun_out/prof_${TAG}_$3.ncu-rep; echo; echo "# top SASS lines by warp-stall samples"; python scripts/ncu_hot.py gpurun_out/prof_${TAG}_$3.ncu-rep 0 20; } > gpurun_out/ncu_full_${TAG}_$3.txt 2>&1
  rm -f gpurun_out/prof_${TAG}_$3.ncu-rep
}
rm -f gpurun_out/*.ncu-rep
cap c2 softmax_ce_reg c2_softmax_ce
cap c4 mixture_loss c4_mixture_loss
cap c2 input_conv_bwd_stage1 c2_input_conv_bwd
