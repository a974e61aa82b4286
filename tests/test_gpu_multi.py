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
