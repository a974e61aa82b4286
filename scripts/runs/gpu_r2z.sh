set -x
export WN_TC_STACK_BWD=2
timeout 600 python -m pytest tests/test_gpu_group_wgrad.py -k "stack_backward" -m gpu -q -x -s > gpurun_out/r2z_t1.log 2>&1; echo "t1 rc=$?"; grep -E "rel-L2|oracle|passed|failed|Error|timed out|assert" gpurun_out/r2z_t1.log | cut -c1-300 | head -12
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not precise and not schedules" > gpurun_out/r2z_t2.log 2>&1; echo "t2 rc=$?"; tail -4 gpurun_out/r2z_t2.log
unset WN_TC_STACK_BWD
for r in 1 2; do for v in 1 2; do echo -n "bwd v$v: "; WN_TC_STACK_BWD=$v python scripts/steptime.py c2 200 | tail -1; done; done
for v in 1 2; do echo -n "c5 bwd v$v: "; WN_TC_STACK_BWD=$v python scripts/steptime.py c5 150 | tail -1; done
for v in 1 2; do echo -n "c3 bwd v$v: "; WN_TC_STACK_BWD=$v python scripts/steptime.py c3 300 | tail -1; done
WN_TC_STACK_BWD=2 WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c2 3 2>&1 | grep -E "SBWD" | tail -5
