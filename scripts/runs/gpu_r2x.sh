set -x
timeout 900 python -m pytest tests/test_gpu_configs.py -k "c3" -m gpu -q -x -s > gpurun_out/r2x_t1.log 2>&1; echo "t1 rc=$?"; grep -E "passed|failed|Error|timed out|assert" gpurun_out/r2x_t1.log | cut -c1-300 | head -8; grep -o '"worst_faithful": [^,]*, "worst_faithful_tensor": "[^"]*"\|"stack_bwd_layers": [0-9]*' gpurun_out/r2x_t1.log | head
timeout 1200 python -m pytest tests -m gpu -q -x -k "not precise" > gpurun_out/r2x_t2.log 2>&1; echo "t2 rc=$?"; tail -4 gpurun_out/r2x_t2.log
scripts/ab.sh c3 300 2 wavenets_b200/libwavenet_b200_head.so wavenets_b200/libwavenet_b200.so
python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2x_c3.json 2> gpurun_out/bench_r2x_c3.err; echo "bench c3 rc=$?"
