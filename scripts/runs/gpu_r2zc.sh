set -x
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2zc_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2zc_tests.log
for c in c2 c1 c3 c4 c5; do
  python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_r2z_$c.json 2> gpurun_out/bench_r2z_$c.err; echo "bench $c rc=$?"; tail -1 gpurun_out/bench_r2z_$c.err
done
python bench.py --config c1 --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2z_c1_bf16.json 2> gpurun_out/bench_r2z_c1_bf16.err; echo "bench c1 bf16 rc=$?"
python bench.py --config c2 --dropout 0.1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2z_c2_dropout.json 2> gpurun_out/bench_r2z_c2_dropout.err; echo "bench c2 dropout rc=$?"
python bench.py --config c3 --dropout 0.1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2z_c3_dropout.json 2> gpurun_out/bench_r2z_c3_dropout.err; echo "bench c3 dropout rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -2
for c in c2 c3; do scripts/ncu_list.sh $c r2z > /dev/null 2>&1; head -6 gpurun_out/launches_r2z_${c}_summary.txt; done
