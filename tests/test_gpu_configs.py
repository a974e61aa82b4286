"""Oracle parity at the FIVE BASELINE.json topologies (`wavenets_b200.config.CONFIGS`: real block counts, widths and
dilation bounds — dilations up to 512, 30-40 blocks, K = L*D skip sum), through the C ABI, on seeded synthetic audio.

Checkers (both independent of the CUDA code; oracle/ is test infrastructure):
  * `oracle/faithful.py`, faithful=False — the reference path in float64 (model.py:213-239,309-335,505-551;
    layers.py:178-224).  fp32 tier: <= 1e-4 relative (north star).  bf16 tier: the STATED bf16 tolerance
    (loss 1e-2, gradients rel-L2 6e-2 per tensor after 30-40 blocks with leaky_relu): this bounds bf16 storage, not the kernels.
  * `oracle/faithful.py`, faithful=True — the same model rounding to bf16 exactly where the kernels do, with the relu /
    leaky_relu derivative masks read back from the device (tests/util.py:device_slope_masks: a pre-activation within one
    rounding flip of zero lands on either side of the kink in two correct implementations; each such element changes a
    gradient five-fold and a handful of them is a percent of a tensor — measured 1.3e-2 without the masks, 2e-3 with).
    What is left is fp32-vs-fp64 accumulation order and MUFU.TANH (2^-11) flipping stored bf16 values by one ulp:
    every gradient tensor within TOL_GRAD_FAITHFUL, all gradients together within TOL_GLOBAL_FAITHFUL, for every
    activation.  A kernel bug of a few percent cannot hide there.  The PRECISE flavour of the library (same sources,
    -DWN_PRECISE_MATH: accurate tanh / sigmoid in the gates) is run on C2 as well: it lands at the same distance (3.4e-3 vs
    3.3e-3), i.e. the residual is accumulation-order flips, not MUFU.TANH.

Shapes: B x T chosen so that a layer has MORE 256-row tiles than the GPU has CTA pairs (74): the persistent stack-forward
launch with its inter-layer flags, the grouped weight gradients with side launches and CUDA-graph replay are all on
the checked path (asserted below).  T > receptive field + 512.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import faithful
from tests import config_parity
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL_FP32 = 1e-4
TOL_LOSS_BF16, TOL_GRAD_BF16 = 1e-2, 6e-2              # shipped library vs float64: stated bf16 tolerance
TOL_GRAD_BF16_NARROW = 1e-1                            # = TOL_GRAD_PWL of tests/test_gpu_parity_bf16.py (relu / leaky_relu models)
TOL_LOSS_FAITHFUL = 5e-5                               # vs the bf16-faithful oracle: loss
TOL_GRAD_FAITHFUL, TOL_GLOBAL_FAITHFUL = 8e-3, 4e-3    # worst gradient tensor / all gradients together (measured: <= 4.7e-3 / 2.1e-3)
TOL_GRAD_FAITHFUL_DROPOUT, TOL_GLOBAL_FAITHFUL_DROPOUT = 1.2e-2, 5e-3   # training pass with dropout 0.1: every block's masked, 1/(1-p)-scaled input
                                                                        # is one more bf16 rounding of x_out (measured C2 4.4e-3 / 2.4e-3, C5's 40 blocks 8.1e-3 / 3.5e-3)
TOL_GRAD_PRECISE, TOL_GLOBAL_PRECISE = 8e-3, 4e-3      # precise flavour (measured on C2: 3.4e-3 / 1.6e-3, shipped 3.3e-3 / 1.7e-3)


def _check_paths(r):
  if r['want_stack']:
    assert r['stack_layers'] == r['blocks'], 'the persistent stack-forward launch must be on the checked path'
    assert r['stack_bwd_layers'] == r['blocks'], 'the persistent stack-backward launch must be on the checked path'
  if r['want_group']:
    assert r['grouped_tiles'] > 0, 'the grouped weight-gradient launch must be on the checked path'


@pytest.mark.parametrize('name', ['c1', 'c1_bf16', 'c2', 'c3', 'c4', 'c5'])
def test_config_topology_vs_oracle(name):
  r = config_parity.run(name)
  print(json.dumps(r))
  _check_paths(r)
  if r['precision'] == 'fp32':
    assert abs(r['loss'] - r['loss_fp64']) <= TOL_FP32 * abs(r['loss_fp64']), r
    assert r['worst_fp64'] <= TOL_FP32, r
    return
  assert abs(r['loss'] - r['loss_fp64']) <= TOL_LOSS_BF16 * abs(r['loss_fp64']), r
  # c1_bf16: 32-entry bias vectors of leaky_relu convs, one sequence — the handful of derivative flips of bf16 storage weighs more in a
  # tensor that small (measured 6.5e-2 on block3/dil1/bias against float64, 3.1e-3 against the bf16-faithful oracle)
  assert r['worst_fp64'] <= (TOL_GRAD_BF16_NARROW if name == 'c1_bf16' else TOL_GRAD_BF16), r
  assert abs(r['loss'] - r['loss_faithful']) <= TOL_LOSS_FAITHFUL * abs(r['loss_faithful']), r
  assert r['worst_faithful'] <= TOL_GRAD_FAITHFUL, r
  assert r['global_faithful'] <= TOL_GLOBAL_FAITHFUL, r


@pytest.mark.parametrize('name', ['c2', 'c3', 'c5'])
def test_config_topology_training_pass_with_dropout(name):
  """defaults.yaml:17 / train.py:22-50 train with dropout 0.1 (layers.py:109-112,192-196).  The keep-masks are applied INSIDE the
  persistent stack launches (forward: the epilogue that produces x_out also writes the next block's masked conv-branch input;
  backward: the OUT epilogue masks the branch gradient before adding the residual gradient), so both stay on the checked path;
  masks injected (TF's RNG stream cannot be reproduced), the same masks in the oracle."""
  r = config_parity.run(name, dropout=0.1)
  print(json.dumps(r))
  _check_paths(r)
  assert r['stack_bwd_layers'] == r['blocks'], 'the persistent stack-backward launch must be on the checked path'
  assert abs(r['loss'] - r['loss_fp64']) <= TOL_LOSS_BF16 * abs(r['loss_fp64']), r
  assert r['worst_fp64'] <= TOL_GRAD_BF16, r
  assert abs(r['loss'] - r['loss_faithful']) <= TOL_LOSS_FAITHFUL * abs(r['loss_faithful']), r
  assert r['worst_faithful'] <= TOL_GRAD_FAITHFUL_DROPOUT, r
  assert r['global_faithful'] <= TOL_GLOBAL_FAITHFUL_DROPOUT, r


@pytest.mark.parametrize('name', ['c2'])
def test_config_topology_precise_flavour_vs_faithful_oracle(name):
  from wavenets_b200 import build
  lib = build.build(flavour='precise')
  env = dict(os.environ, WN_LIB=lib, PYTHONPATH=ROOT)
  p = subprocess.run([sys.executable, '-m', 'tests.config_parity', name], capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
  assert p.returncode == 0, p.stderr[-3000:]
  r = json.loads([l for l in p.stdout.splitlines() if l.startswith('{')][-1])
  print(json.dumps(r))
  _check_paths(r)
  assert abs(r['loss'] - r['loss_faithful']) <= TOL_LOSS_FAITHFUL * abs(r['loss_faithful']), r
  assert r['worst_faithful'] <= TOL_GRAD_PRECISE, r
  assert r['global_faithful'] <= TOL_GLOBAL_PRECISE, r


def test_c2_forward_vs_faithful():
  """WaveNet.call at the C2 topology: probabilities against the oracles (and softmax rows sum to one)."""
  m, ocfg, p, x, cond, precision, _, _, _ = config_parity.setup('c2')
  y = m((x[:, :-1], cond)).cpu().numpy()
  ref = faithful.forward(p, ocfg, x[:, :-1], cond, faithful=True)
  ref64 = faithful.forward(p, ocfg, x[:, :-1], cond, faithful=False)
  e_f, e_64 = rel_l2(y, ref), rel_l2(y, ref64)
  print(f'c2 forward: rel-L2 vs bf16-faithful {e_f:.2e}, vs fp64 {e_64:.2e}')
  assert e_f <= 1e-2 and e_64 <= 1e-2
  assert np.allclose(y.sum(-1), 1.0, atol=1e-4)
