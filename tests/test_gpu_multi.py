"""Data-parallel correctness ON HARDWARE (needs >= 2 GPUs on the box; skipped otherwise): torchrun with 2 ranks, NCCL.
The all-reduced gradients of two batch shards (loss divided by the GLOBAL batch, model.py:328; SUM all-reduce standing in for
MirroredStrategy, train.py:203, model.py:336) equal the 1-GPU gradients of the concatenated batch, fp32 tier <= 1e-5 relative
(reduction order only), through BOTH collectives: ncclAllReduce behind the C ABI (wn_allreduce_grads, inside the step graph)
and torch.distributed.all_reduce on the flat buffer."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


@pytest.mark.parametrize('native', [1, 0])
def test_two_gpu_allreduced_grads_equal_single_gpu(native):
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs on one box (gpurun --gpus 2)')
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
         '--master-port', str(_free_port()), os.path.join(ROOT, 'scripts', 'dp_grad_check.py'), '--native', str(native)]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
  lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith('{')]
  assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
  assert len(lines) == 4 and all(l['ok'] for l in lines), lines
  if native:
    assert all('C ABI' in l['allreduce'] for l in lines)
    # the bucketed all-reduce (first group of blocks reduced beside the second group's weight-gradient kernels) is on the checked path
    bk = [l for l in lines if l['case'] == 'bf16_stack_buckets'][0]
    assert bk['stack_backward_layers'] == 6 and bk['early_allreduce_buckets'] >= 1, bk


def test_two_devices_in_one_process():
  """One process, handles on two GPUs (the reference drives all replicas from one Python process, train.py:203-205): the launchers'
  one-time setup (dynamic shared memory attribute, cluster occupancy) is per device.  Same weights and inputs on cuda:0 and cuda:1 —
  the persistent stack launches, the grouped weight gradients and CUDA-graph replay included — give bit-identical results."""
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs on one box (gpurun --gpus 2)')
  import numpy as np
  from tests.util import make_inputs
  from wavenets_b200 import WaveNet
  kw = dict(channels=256, blocks=3, layers_per_block=2, dilation_bound=16, skip_channels=256, final_layers_channels=[128], activation='leaky_relu')
  B, T = 3, 6500
  x, _ = make_inputs(B, T, 0)
  res = []
  weights = None
  for dev in (0, 1):
    m = WaveNet(**kw, precision='bf16', device=dev)
    m.build(x[:, :-1].shape)
    if weights is None:
      m.handle.glorot_init(seed=4, bias_std=0.02)
      weights = m.get_weights()
    else:
      m.set_weights(weights)
    for _ in range(3):
      out = m.train_step(x)
    res.append((out['loss'], m.get_grads(), int(m.handle.lib.wn_stack_backward_layers(m.handle.h))))
  assert res[0][2] == 3 and res[1][2] == 3
  assert res[0][0] == res[1][0]
  for k in res[0][1]:
    assert np.array_equal(res[0][1][k], res[1][1][k]), k
