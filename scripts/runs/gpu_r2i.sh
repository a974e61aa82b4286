set -x
python -m pytest tests/test_gpu_group_wgrad.py -k "stack_forward" tests/test_gpu_configs.py -m gpu -q -x -s > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; grep -E "rel-L2|passed|failed|Error|assert" gpurun_out/r2i_tests.log | cut -c1-300 | head -20
WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c3 3 2>&1 | grep -E "SFWD|loss" > gpurun_out/r2i_timeline_c3.log; tail -5 gpurun_out/r2i_timeline_c3.log
WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c2 3 2>&1 | grep -E "SFWD|loss" > gpurun_out/r2i_timeline_c2.log; tail -5 gpurun_out/r2i_timeline_c2.log
