set -x
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/final_tests.log
python bench.py --config c1 --steps 50 --warmup 5 > gpurun_out/bench_final_c1.json 2> gpurun_out/bench_final_c1.err; echo "bench c1 rc=$?"
python bench.py > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final_default.err; echo "bench default rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -4
