for c in c1 c3 c4 c5; do
  echo "== $c"
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
s=sys.stdin.read()
try:
  d=json.loads(s); print(d['config']['workload'], '|', round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms', d['dtype'], 'dil TF', round(d['roofline']['achieved'],1), 'whole TF', round(d['whole_step']['tflops_all_gemms_whole_step'],1), 'loss', d['loss'])
except Exception as e:
  print('FAILED', s[-1500:])
"
done
