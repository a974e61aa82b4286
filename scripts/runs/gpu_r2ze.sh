set -x
python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2ze_multi.log 2>&1; echo "multi rc=$?"; tail -12 gpurun_out/r2ze_multi.log | cut -c1-300
