set -x
python -m pytest tests/test_gpu_group_wgrad.py -k "stack_forward" tests/test_gpu_configs.py -m gpu -q -x -s > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; grep -E "rel-L2|passed|failed|Error|assert" gpurun_out/r2h_tests.log | cut -c1-300 | head -20
python -m pytest tests -m gpu -q -x > gpurun_out/r2h_tests_all.log 2>&1; echo "all tests rc=$?"; tail -5 gpurun_out/r2h_tests_all.log
scripts/ab.sh c3 300 2 wavenets_b200/libwavenet_b200_base.so wavenets_b200/libwavenet_b200.so
python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2h_c3.json 2> gpurun_out/bench_r2h_c3.err; echo "bench c3 rc=$?"; tail -2 gpurun_out/bench_r2h_c3.err
