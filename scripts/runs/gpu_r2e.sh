set -x
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -s > gpurun_out/r2e_fullsize.log 2>&1; echo "fullsize rc=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/r2e_fullsize.log | head
WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c2 3 2>&1 | grep -E "SBWD|loss" > gpurun_out/r2e_timeline_c2.log; tail -12 gpurun_out/r2e_timeline_c2.log
ncu --set full --clock-control none --import-source on -k "regex:tc_stack_bwd" --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_r2e_stackbwd python scripts/prof_step.py c2 3 > gpurun_out/ncu_r2e_a.log 2>&1; tail -2 gpurun_out/ncu_r2e_a.log
