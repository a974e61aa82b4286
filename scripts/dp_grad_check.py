"""Multi-GPU correctness of the data-parallel path (SURVEY.md 8e "scaling check"; train.py:203, model.py:328,336): the
all-reduced gradients of N batch shards, each computed with the loss divided by the GLOBAL batch, equal the gradients of
ONE GPU running the concatenated batch.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 --master-port P scripts/dp_grad_check.py [--native 0|1]; rank 0 prints one JSON line per case and
exits non-zero when a case misses its tolerance."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from wavenets_b200 import WaveNet, parallel, synth

CASES = {
  # fp32 tier: reduction-order differences only -> <= 1e-5 relative
  'fp32_cond_skip': (dict(channels=32, blocks=4, layers_per_block=2, activation='leaky_relu', dilation_bound=16, skip_channels=48,
                          final_layers_channels=[64, 64], conditioning='global', mapping_layers=[8, 16], mapping_activation='leaky_relu'), 'fp32', 2, 1500, 1e-5),
  # bf16 tier, grouped weight gradients + fused forward: sequences are independent, so the shards' roundings are the 1-GPU run's
  'bf16_r256': (dict(channels=256, blocks=4, layers_per_block=1, dilation_bound=16, skip_channels=256, final_layers_channels=[256],
                     conditioning='global', mapping_layers=[8], mapping_activation='tanh'), 'bf16', 2, 2048, 2e-5),
  # the persistent stack launches (75 row tiles per replica > 74 CTA pairs) with the BUCKETED all-reduce: the grouped weight-gradient
  # launch runs in two groups of blocks, the first group's slice of the flat gradient buffer is reduced beside the second group's kernels
  'bf16_stack_buckets': (dict(channels=256, blocks=6, layers_per_block=1, dilation_bound=16, skip_channels=256, final_layers_channels=[256],
                              conditioning='global', mapping_layers=[8], mapping_activation='tanh'), 'bf16', 3, 6400, 3e-5),
  # dropout with injected keep-masks: every replica gets its rows of the global masks (Philox masks differ per replica by design)
  'fp32_dropout_masks': (dict(channels=32, blocks=3, layers_per_block=1, dilation_bound=8, final_layers_channels=[32], dropout=0.25), 'fp32', 2, 600, 1e-5),
}


def run_case(name, rank, local, world, native):
  kw, precision, b_local, T, tol = CASES[name]
  cond_in = 11 if kw.get('conditioning') else 0
  B = b_local * world
  dev = torch.device('cuda', local)
  x = synth.frames(B, T, seed=7)
  cond = synth.speakers_onehot(B, cond_in, seed=7) if cond_in else None
  masks = None
  if kw.get('dropout', 0) > 0:
    rng = np.random.default_rng(11)
    masks = [rng.random((B, T, kw['channels'])) >= kw['dropout'] for _ in range(kw['blocks'])]
  lo, hi = parallel.shard_bounds(B, rank, world)
  m = WaveNet(**kw, precision=precision, device=local, max_batch=b_local, max_time=T)
  m.build(((b_local, T, 1), (b_local, cond_in)) if cond_in else (b_local, T, 1))
  m.handle.glorot_init(seed=1, bias_std=0.02)
  parallel.attach(m, native=native)
  if masks is not None:
    m.set_dropout_masks([k[lo:hi] for k in masks])
  data = (x[lo:hi], cond[lo:hi]) if cond is not None else x[lo:hi]
  for _ in range(3):                       # eager, plans, CUDA-graph replay (with the all-reduce inside when native)
    loss = m.train_step_async(data)
  g_dp = m.handle.flat_grads.clone()
  loss_dp = loss[0:1].clone()
  dist.all_reduce(loss_dp)                 # every replica reports its own share of the global-batch mean (model.py:328)
  res = None
  if rank == 0:
    m1 = WaveNet(**kw, precision=precision, device=local, max_batch=B, max_time=T)
    m1.build(((B, T, 1), (B, cond_in)) if cond_in else (B, T, 1))
    m1.set_weights(m.get_weights())
    if masks is not None:
      m1.set_dropout_masks(masks)
    l1 = m1.train_step((x, cond) if cond is not None else x)['loss']
    g1 = m1.handle.flat_grads
    worst, who = 0.0, None
    hh = m.handle
    for nm, shape, off in zip(hh.names, hh.shapes, hh.offsets):
      n = int(np.prod(shape))
      a, b = g_dp[off:off + n].double(), g1[off:off + n].double()
      if float(b.norm()) == 0.0:
        assert float(a.abs().max()) == 0.0, nm
        continue
      e = float((a - b).norm() / b.norm())
      if e > worst:
        worst, who = e, nm
    res = {'case': name, 'n_replicas': world, 'precision': precision, 'global_batch': B, 'T': T, 'worst_rel_l2': worst, 'tensor': who, 'tol': tol,
           'loss_sum_of_replicas': float(loss_dp.item()), 'loss_single_gpu': l1, 'allreduce': 'C ABI (wn_allreduce_grads, NCCL)' if m._comm is not None else 'torch.distributed',
           'early_allreduce_buckets': int(hh.lib.wn_allreduce_buckets(hh.h)), 'stack_backward_layers': int(hh.lib.wn_stack_backward_layers(hh.h)),
           'ok': bool(worst <= tol and abs(float(loss_dp.item()) - l1) <= 1e-5 * abs(l1))}
    print(json.dumps(res), flush=True)
  dist.barrier()
  return res


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--native', type=int, default=1)
  ap.add_argument('--cases', default=','.join(CASES))
  args = ap.parse_args()
  rank, local, world = parallel.init_from_env()
  torch.cuda.set_device(local)
  ok = True
  for name in args.cases.split(','):
    r = run_case(name, rank, local, world, bool(args.native))
    if rank == 0 and not r['ok']:
      ok = False
  dist.barrier()
  dist.destroy_process_group()
  sys.exit(0 if ok else 1)


if __name__ == '__main__':
  main()
