set -x
python -m pytest tests/test_gpu_parity_fp32.py tests/test_gpu_golden.py tests/test_gpu_optimizer.py tests/test_gpu_configs.py -k "not precise and not c2 and not c4 and not c5 and not c3" -m gpu -q -x > gpurun_out/r2zf_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2zf_tests.log
scripts/ab.sh c1 500 2 wavenets_b200/libwavenet_b200_head.so wavenets_b200/libwavenet_b200.so
python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2zf_c1.json 2> gpurun_out/bench_r2zf_c1.err; echo "bench c1 rc=$?"
