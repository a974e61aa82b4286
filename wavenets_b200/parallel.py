"""Data parallelism: the reference's only strategy (tf.distribute.MirroredStrategy, train.py:203).

One process per GPU.  Rank r trains on sequences [r*B/N, (r+1)*B/N) of the global batch; the
loss (and therefore every gradient) is already divided by the GLOBAL batch
(tf.nn.compute_average_loss, model.py:328), so replicas' gradients are SUM-reduced: one
all-reduce over the flat fp32 gradient buffer per step (NCCL over NVLink on GPUs; gloo in the
CPU tests of this host logic).  No other collective exists on the path.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
  """Initialise torch.distributed from torchrun's env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).
  Returns (rank, local_rank, world_size); a no-op returning (0, 0, 1) outside torchrun."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1 and not dist.is_initialized():
    if backend is None:
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    if backend == 'nccl':
      torch.cuda.set_device(local)
      dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device('cuda', local))
    else:
      dist.init_process_group(backend, rank=rank, world_size=world)
  return rank, local, world


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
  """Rows [lo, hi) of the global batch owned by `rank` (equal shards, like MirroredStrategy)."""
  if global_batch % world != 0:
    raise ValueError(f'global batch {global_batch} is not divisible by {world} replicas')
  per = global_batch // world
  return rank * per, (rank + 1) * per


def shard_batch(x, rank: int, world: int):
  """Slice a (B, ...) array/tensor (or a tuple of them) to this rank's sequences."""
  if isinstance(x, (tuple, list)):
    return tuple(shard_batch(t, rank, world) for t in x)
  lo, hi = shard_bounds(int(x.shape[0]), rank, world)
  return x[lo:hi]


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
  """In-place SUM all-reduce of the flat gradient buffer (the only collective of the path)."""
  if dist.is_initialized() and dist.get_world_size(group) > 1:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  return flat


def attach(model, group=None, native: Optional[bool] = None):
  """Make `model.train_step` behave like a MirroredStrategy replica: losses are divided by
  B_local * n_replicas and gradients are SUM-all-reduced after the backward pass.

  On GPUs the all-reduce runs behind the C ABI (`wn_comm_init` / `wn_allreduce_grads`): rank 0 draws an NCCL unique id in
  the C library, torch.distributed only carries those 128 bytes to the other ranks.  `native=False` (and every CPU / gloo
  process group) keeps `torch.distributed.all_reduce` on the flat gradient buffer instead."""
  world = dist.get_world_size(group) if dist.is_initialized() else 1
  rank = dist.get_rank(group) if dist.is_initialized() else 0
  model.n_replicas = world
  model.replica_rank = rank
  model._process_group = None
  model._comm = None
  if world > 1:
    model._process_group = group if group is not None else dist.group.WORLD
    if native is None:
      native = torch.cuda.is_available() and dist.get_backend(group) == 'nccl'
    if native:
      from . import _lib
      lib = _lib.load()
      idbuf = torch.zeros(128, dtype=torch.uint8)
      if rank == 0:
        import ctypes as C
        raw = (C.c_uint8 * 128)()
        _lib.check(lib.wn_nccl_unique_id(raw))
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
      dev = torch.device('cuda', torch.cuda.current_device())
      idbuf = idbuf.to(dev)
      dist.broadcast(idbuf, src=dist.get_global_rank(model._process_group, 0) if hasattr(dist, 'get_global_rank') else 0,
                     group=model._process_group)
      model._comm = (bytes(idbuf.cpu().tolist()), world, rank)
      if getattr(model, '_handle', None) is not None:
        model._init_comm(*model._comm)
  if getattr(model, '_handle', None) is not None and getattr(model, 'dropout', 0) > 0:
    model.set_dropout_seed(model._dropout_seed)     # re-key with this replica's rank
  return model


def replica_loss_scale(b_local: int, n_replicas: int) -> float:
  """tf.nn.compute_average_loss: sum over everything / (B_local * n_replicas)."""
  return 1.0 / (float(b_local) * float(n_replicas))
