#!/bin/bash
# ncu --set full captures (one GPU) of the GEMM-class kernels of a C2 step, forward and backward.
# usage (on the GPU box): scripts/ncu_capture.sh <tag>   -> gpurun_out/prof_<tag>_{fwd,bwd}.ncu-rep
set -e
TAG=${1:-r1c}
python scripts/prof_step.py c2 3 > gpurun_out/plain_$TAG.log 2>&1     # must exit 0 without ncu first
K='regex:tc_conv_gemm_staged|tc_wgrad'
# step layout (matching launches): 63 forward (gate, conv1 x30 + skip/head), then per block: wgrad 1x1, finish, gate_bwd, wgrad dilated, finish, dgrad
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 200 --launch-count 4 -f -o gpurun_out/prof_${TAG}_fwd python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 282 --launch-count 8 -f -o gpurun_out/prof_${TAG}_bwd python scripts/prof_step.py c2 3 > gpurun_out/ncu_${TAG}_bwd.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_fwd.log gpurun_out/ncu_${TAG}_bwd.log
