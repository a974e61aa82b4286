set -x
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_configs.py -k "fullsize or c2 or c5 or c4" tests/test_gpu_group_wgrad.py -m gpu -q -x > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2r_tests.log
scripts/ab.sh c2 200 2 wavenets_b200/libwavenet_b200_head.so wavenets_b200/libwavenet_b200.so
scripts/ab.sh c5 150 1 wavenets_b200/libwavenet_b200_head.so wavenets_b200/libwavenet_b200.so
WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c2 3 2>&1 | grep -E "SBWD" | tail -5
scripts/ncu_list.sh c2 r2r > /dev/null 2>&1; head -5 gpurun_out/launches_r2r_c2_summary.txt
