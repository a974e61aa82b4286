"""CPU tests (gloo, world_size 2) of the data-parallel host logic: batch sharding, the
replica loss scale of tf.nn.compute_average_loss (model.py:328) and the SUM all-reduce of the flat
gradient buffer.  The per-replica "local step" is the CPU oracle (the CUDA path needs a GPU): the
property checked is the one the GPU path relies on — all-reduced per-shard gradients, each
computed with scale 1/(B_local * n_replicas), equal the single-replica gradients of the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wavenet_oracle as wo
from tests.util import SMALL_MODELS, COND_IN, make_inputs, oracle_config


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, name, B, T, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  from wavenets_b200 import parallel
  r, _, w = parallel.init_from_env(backend='gloo')
  assert (r, w) == (rank, world)
  kw = SMALL_MODELS[name]
  cond_in = COND_IN if kw.get('conditioning') else 0
  cfg = oracle_config(kw, cond_in)
  p = wo.init_params(cfg, seed=1)
  x, cond = make_inputs(B, T, cond_in)
  xs = parallel.shard_batch(x, rank, world)
  cs = parallel.shard_batch(cond, rank, world) if cond is not None else None
  lo, hi = parallel.shard_bounds(B, rank, world)
  assert xs.shape[0] == hi - lo == B // world
  loss, g, _ = wo.train_step(p, cfg, xs.astype(np.float64), None if cs is None else cs.astype(np.float64), n_replicas=world)
  names = sorted(g)
  flat = torch.from_numpy(np.concatenate([np.asarray(g[k], np.float64).ravel() for k in names]))
  parallel.allreduce_sum_(flat)
  lt = torch.tensor([loss], dtype=torch.float64)
  parallel.allreduce_sum_(lt)
  if rank == 0:
    np.save(os.path.join(out_dir, 'flat.npy'), flat.numpy())
    np.save(os.path.join(out_dir, 'loss.npy'), lt.numpy())
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('name', ['cond_skip', 'categorical_multidil'])
def test_allreduced_shard_grads_equal_full_batch_grads(name, tmp_path):
  B, T, world = 4, 40, 2
  port = _free_port()
  mp.spawn(_worker, args=(world, port, name, B, T, str(tmp_path)), nprocs=world, join=True)
  kw = SMALL_MODELS[name]
  cond_in = COND_IN if kw.get('conditioning') else 0
  cfg = oracle_config(kw, cond_in)
  p = wo.init_params(cfg, seed=1)
  x, cond = make_inputs(B, T, cond_in)
  loss, g, _ = wo.train_step(p, cfg, x.astype(np.float64), None if cond is None else cond.astype(np.float64))
  ref = np.concatenate([np.asarray(g[k], np.float64).ravel() for k in sorted(g)])
  got = np.load(tmp_path / 'flat.npy')
  assert np.abs(got - ref).max() <= 1e-10 * (np.abs(ref).max() + 1e-30)
  assert abs(float(np.load(tmp_path / 'loss.npy')[0]) - loss) <= 1e-10 * abs(loss)


def test_shard_bounds_and_errors():
  from wavenets_b200 import parallel
  assert [parallel.shard_bounds(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
  with pytest.raises(ValueError):
    parallel.shard_bounds(7, 0, 2)
  assert parallel.replica_loss_scale(8, 8) == 1.0 / 64.0
  # outside torchrun: world size 1, attach is a no-op
  class M:
    pass
  m = parallel.attach(M())
  assert m.n_replicas == 1 and m._process_group is None
