set -x
python -m pytest tests/test_gpu_configs.py tests/test_gpu_fullsize.py tests/test_gpu_parity_bf16.py tests/test_gpu_golden.py tests/test_gpu_group_wgrad.py -m gpu -q -x -s > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/r2f_tests.log | cut -c1-300 | head -20
python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2f_c2.json 2> gpurun_out/bench_r2f_c2.err; echo "bench c2 rc=$?"; tail -2 gpurun_out/bench_r2f_c2.err
python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2f_c5.json 2> gpurun_out/bench_r2f_c5.err; echo "bench c5 rc=$?"
WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c2 3 2>&1 | grep -E "SBWD|loss" > gpurun_out/r2f_timeline_c2.log; tail -6 gpurun_out/r2f_timeline_c2.log
scripts/ncu_list.sh c2 r2f
