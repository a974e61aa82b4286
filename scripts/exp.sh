for X in "-DTC_EXP_NO_IN" "-DTC_EXP_NO_OUT" "-DTC_EXP_NO_IN -DTC_EXP_NO_OUT"; do
  echo "== $X"
  WN_NVCC_EXTRA="$X" python -m wavenets_b200.build > /dev/null 2>&1
  WN_NVCC_EXTRA="$X" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_step_in_kernel'], d['whole_step']['gemm_ms_per_step'])"
done
