set -x
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2k_tests.log
scripts/ab.sh c3 300 2 wavenets_b200/libwavenet_b200_base.so wavenets_b200/libwavenet_b200.so
scripts/ab.sh c2 200 2 wavenets_b200/libwavenet_b200_base.so wavenets_b200/libwavenet_b200.so
scripts/ab.sh c1 500 1 wavenets_b200/libwavenet_b200_base.so wavenets_b200/libwavenet_b200.so
WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c3 3 2>&1 | grep -E "SFWD|loss" | tail -4
WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c2 3 2>&1 | grep -E "SFWD|loss" | tail -4
python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2k_c3.json 2> gpurun_out/bench_r2k_c3.err; echo "bench c3 rc=$?"
