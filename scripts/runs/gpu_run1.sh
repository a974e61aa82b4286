set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python -c "import wavenets_b200._lib as l; print(l.load().wn_nccl_info().decode())"
python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_configs.py > gpurun_out/r2a_tests_old.log 2>&1; echo "old tests rc=$?"; tail -15 gpurun_out/r2a_tests_old.log
python -m pytest tests/test_gpu_configs.py -m gpu -q -s > gpurun_out/r2a_tests_cfg.log 2>&1; echo "cfg tests rc=$?"; grep -E "^c[1-5]|passed|failed|Error|assert" gpurun_out/r2a_tests_cfg.log | cut -c1-600 | head -40
for c in c2 c1 c3 c4 c5; do
  python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_r2a_$c.json 2> gpurun_out/bench_r2a_$c.err; echo "bench $c rc=$?"; tail -3 gpurun_out/bench_r2a_$c.err
done
python __graft_entry__.py smoke 2>&1 | tail -5
scripts/ncu_list.sh c2 r2a
