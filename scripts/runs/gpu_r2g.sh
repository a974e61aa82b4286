set -x
python -m pytest tests/test_gpu_configs.py -k "dropout" tests/test_gpu_group_wgrad.py tests/test_gpu_golden.py -m gpu -q -x -s > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; grep -E "rel-L2|passed|failed|Error|\"config\"" gpurun_out/r2g_tests.log | cut -c1-700 | head -20
scripts/ab.sh c2 200 2 wavenets_b200/libwavenet_b200_base.so wavenets_b200/libwavenet_b200.so
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline --dropout 0.1 > gpurun_out/bench_r2g_c2_drop.json 2> gpurun_out/bench_r2g_c2_drop.err; echo "bench c2 dropout rc=$?"; tail -2 gpurun_out/bench_r2g_c2_drop.err
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2g_c2.json 2> gpurun_out/bench_r2g_c2.err; echo "bench c2 rc=$?"
