set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s > gpurun_out/r2n_multi.log 2>&1; echo "multi rc=$?"; grep -E "passed|failed|Error|worst_rel" gpurun_out/r2n_multi.log | cut -c1-600 | head
for nb in 1 2 3; do
WN_AR_BUCKETS=$nb python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --config c2 --check-grads > gpurun_out/bench_r2n_c2_2gpu_nb$nb.json 2> gpurun_out/bench_r2n_c2_2gpu_nb$nb.err; echo "bench nb=$nb rc=$?"; tail -2 gpurun_out/bench_r2n_c2_2gpu_nb$nb.err
done
python bench.py --config c2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2n_c2_1gpu.json 2> gpurun_out/bench_r2n_c2_1gpu.err; echo "bench 1gpu rc=$?"
