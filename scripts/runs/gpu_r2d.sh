set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2d_tests.log
python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/bench_r2d_c2.json 2> gpurun_out/bench_r2d_c2.err; echo "bench c2 rc=$?"; tail -2 gpurun_out/bench_r2d_c2.err
WN_TC_STACK_BWD=0 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2d_c2_nostackbwd.json 2> gpurun_out/bench_r2d_c2_nsb.err; echo "bench c2 nsb rc=$?"
python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2d_c5.json 2> gpurun_out/bench_r2d_c5.err; echo "bench c5 rc=$?"
scripts/ncu_list.sh c2 r2d
