set -x
timeout 900 python -m pytest tests/test_gpu_configs.py -k "dropout" -m gpu -q -x -s > gpurun_out/r2y_t1.log 2>&1; echo "t1 rc=$?"; grep -E "passed|failed|Error|timed out|assert" gpurun_out/r2y_t1.log | cut -c1-300 | head -8; grep -o '"config": "[^"]*"\|"worst_faithful": [^,]*, "worst_faithful_tensor": "[^"]*"\|"stack_bwd_layers": [0-9]*' gpurun_out/r2y_t1.log | head -12
python bench.py --config c3 --dropout 0.1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2y_c3_dropout.json 2> gpurun_out/bench_r2y_c3_dropout.err; echo "bench c3 dropout rc=$?"
