set -x
python -m pytest tests/test_gpu_parity_bf16.py -k "narrow32" -m gpu -q -x > gpurun_out/r2o_narrow.log 2>&1; echo "narrow rc=$?"; tail -25 gpurun_out/r2o_narrow.log | cut -c1-300
python -m pytest tests/test_gpu_configs.py -k "c1_bf16" -m gpu -q -x -s > gpurun_out/r2o_c1bf16.log 2>&1; echo "c1_bf16 rc=$?"; tail -5 gpurun_out/r2o_c1bf16.log | cut -c1-1200
python bench.py --config c1 --precision bf16 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2o_c1_bf16.json 2> gpurun_out/bench_r2o_c1_bf16.err; echo "bench c1 bf16 rc=$?"; tail -3 gpurun_out/bench_r2o_c1_bf16.err
python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2o_c1.json 2> gpurun_out/bench_r2o_c1.err; echo "bench c1 rc=$?"
