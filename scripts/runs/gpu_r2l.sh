set -x
python -m pytest tests/test_gpu_group_wgrad.py tests/test_gpu_configs.py tests/test_gpu_parity_bf16.py -m gpu -q -x > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log
scripts/ab.sh c3 300 2 wavenets_b200/libwavenet_b200_base.so wavenets_b200/libwavenet_b200.so
python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2l_c3.json 2> gpurun_out/bench_r2l_c3.err; echo "bench c3 rc=$?"
