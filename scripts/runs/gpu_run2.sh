set -x
python -m pytest tests -m gpu -q --deselect tests/test_gpu_configs.py > gpurun_out/r2c_tests_old.log 2>&1; echo "old tests rc=$?"; tail -12 gpurun_out/r2c_tests_old.log
python -m pytest tests/test_gpu_configs.py -m gpu -q -s > gpurun_out/r2c_tests_cfg.log 2>&1; echo "cfg tests rc=$?"; grep -E 'passed|failed|c2 forward' gpurun_out/r2c_tests_cfg.log | cut -c1-300
