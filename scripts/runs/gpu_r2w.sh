set -x
timeout 600 python -m pytest tests/test_gpu_group_wgrad.py -k "multi_dilation_vs_per_block_chain" -m gpu -q -x -s > gpurun_out/r2w_t1.log 2>&1; echo "t1 rc=$?"; grep -E "rel-L2|oracle|passed|failed|Error|timed out|assert" gpurun_out/r2w_t1.log | cut -c1-300 | head -12
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_group_wgrad.py -m gpu -q -x -k "not precise" > gpurun_out/r2w_t2.log 2>&1; echo "t2 rc=$?"; tail -4 gpurun_out/r2w_t2.log
