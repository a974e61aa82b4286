set -x
for r in 1 2; do for v in 1 2; do for x in 0 1; do echo -n "bwd v$v relaxed $x: "; WN_TC_HANDOFF_RELAXED=$x WN_TC_STACK_BWD=$v python scripts/steptime.py c2 200 | tail -1; done; done; done
WN_TC_HANDOFF_RELAXED=1 WN_TC_STACK_BWD=1 WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c2 3 2>&1 | grep -E "SBWD" | tail -3
export WN_TC_HANDOFF_RELAXED=1
timeout 900 python -m pytest tests/test_gpu_group_wgrad.py -k "stack_backward" tests/test_gpu_configs.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not precise" > gpurun_out/r2zb_t1.log 2>&1; echo "t1 rc=$?"; tail -3 gpurun_out/r2zb_t1.log
