set -x
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_configs.py -k "fullsize or c2 or c5" tests/test_gpu_group_wgrad.py -m gpu -q -x > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2q_tests.log
for r in 1 2; do for b in 0 128 96; do echo -n "band $b: "; WN_TC_SB_BAND=$b python scripts/steptime.py c2 200 | tail -1; done; done
for b in 0 128; do echo -n "c5 band $b: "; WN_TC_SB_BAND=$b python scripts/steptime.py c5 200 | tail -1; done
WN_TC_SB_BAND=128 scripts/ncu_list.sh c2 r2q > /dev/null 2>&1; head -5 gpurun_out/launches_r2q_c2_summary.txt
