set -x
python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity_fp32.py tests/test_gpu_parity_bf16.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -k "not precise" -m gpu -q -x > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2v_tests.log
scripts/ab.sh c2 200 2 wavenets_b200/libwavenet_b200_head.so wavenets_b200/libwavenet_b200.so
python bench.py --config c2 --dropout 0.1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2v_c2_dropout.json 2> gpurun_out/bench_r2v_c2_dropout.err; echo "bench c2 dropout rc=$?"
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2v_c2.json 2> gpurun_out/bench_r2v_c2.err; echo "bench c2 rc=$?"
