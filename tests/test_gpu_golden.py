"""GPU parity against the golden vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/,
written by oracle/make_golden.py).  The CUDA path is called through the C ABI with the fixture's
inputs and weights; no oracle in the loop.  Tolerances: fp32 tier <= 1e-4 relative (north star);
bf16 tier as stated in tests/test_gpu_parity_bf16.py."""
import numpy as np
import pytest

from tests.golden_util import CASES, load_case
from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4
TOL_OUT_BF16, TOL_LOSS_BF16, TOL_GRAD_BF16_PWL, MIN_COS_PWL = 1e-2, 1e-2, 1e-1, 0.995


def _build(c, precision):
  from wavenets_b200 import WaveNet
  m = WaveNet(**c.kw, precision=precision)
  x_in = c.x[:, :-1]
  m.build((x_in.shape, c.cond.shape) if c.cond is not None else x_in.shape)
  assert m.variable_names == list(c.weights)
  assert m.receptive_field == c.receptive_field
  m.n_replicas = c.n_replicas
  m.set_weights(c.weights)
  return m


@pytest.mark.parametrize('name', sorted(CASES))
def test_fp32_tier_matches_reference_golden(name):
  c = load_case(name)
  m = _build(c, 'fp32')
  data = (c.x, c.cond) if c.cond is not None else c.x
  x_in = c.x[:, :-1]
  # WaveNet.call (inference mode)
  pred = m((x_in, c.cond) if c.cond is not None else x_in).cpu().numpy()
  assert rel_err(pred[:, c.pred_t], c.pred) < TOL
  # prepare_target: bit-exact
  if c.cfg.num_mixtures is None:
    assert np.array_equal(m.prepare_target(c.x[:, 1:, :]).cpu().numpy(), c.target)
  # test_step
  assert abs(m.test_step(data)['loss'] - c.test_loss) <= TOL * abs(c.test_loss)
  # train_step: metrics + every gradient handed to optimizer.apply_gradients
  if c.keep_masks is not None:
    m.set_dropout_masks(c.keep_masks)
  out = m.train_step(data)
  assert abs(out['loss'] - c.train_loss) <= TOL * abs(c.train_loss), (out['loss'], c.train_loss)
  if c.reg_loss is not None:
    assert abs(out['reg_loss'] - c.reg_loss) <= TOL * abs(c.reg_loss)
  g = m.get_grads()
  for k, ref in c.grads.items():
    assert rel_err(g[k], ref) < TOL, (k, rel_err(g[k], ref))


@pytest.mark.parametrize('name', ['tc_cond_skip64', 'tc_multidil64'])
def test_bf16_tier_matches_reference_golden(name):
  c = load_case(name)
  m = _build(c, 'bf16')
  data = (c.x, c.cond) if c.cond is not None else c.x
  x_in = c.x[:, :-1]
  pred = m((x_in, c.cond) if c.cond is not None else x_in).cpu().numpy()
  assert rel_l2(pred[:, c.pred_t], c.pred) < TOL_OUT_BF16
  assert abs(m.test_step(data)['loss'] - c.test_loss) <= TOL_LOSS_BF16 * abs(c.test_loss)
  out = m.train_step(data)
  assert abs(out['loss'] - c.train_loss) <= TOL_LOSS_BF16 * abs(c.train_loss)
  g = m.get_grads()
  for k, ref in c.grads.items():
    if np.linalg.norm(ref) == 0:
      continue
    a, b = g[k].ravel().astype(np.float64), ref.ravel()
    assert rel_l2(a, b) < TOL_GRAD_BF16_PWL, (k, rel_l2(a, b))         # leaky_relu models: see test_gpu_parity_bf16.py header
    assert float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b))) > MIN_COS_PWL, k


@pytest.mark.parametrize('name', ['cat_multidil', 'cond_skip', 'tc_cond_skip64'])
def test_layer_call_matches_reference_golden(name):
  import torch
  from wavenets_b200 import WaveNetLayer
  c = load_case(name)
  kw = c.kw
  lay = WaveNetLayer(kernel=kw.get('kernel_size', 2), dilation_rate=[int(d) for d in c.dilations[0]], activation=kw.get('activation'),
                     channels=kw['channels'], residual=kw.get('use_residual', True), dilation_channels=kw.get('dilation_channels'),
                     skip_channels=kw.get('skip_channels'), condition=c.layer0_cond is not None)
  x = c.layer0_x
  lay.build((x.shape, (x.shape[0], x.shape[1], c.layer0_cond.shape[-1])) if c.layer0_cond is not None else x.shape)
  lay.set_weights({k[len('block0/'):]: v for k, v in c.weights.items() if k.startswith('block0/')})
  if c.layer0_cond is not None:
    cond_t = np.repeat(c.layer0_cond[:, None, :], x.shape[1], axis=1)
    xo, sk = lay.call((x, cond_t))
  else:
    xo, sk = lay.call(x)
  assert rel_err(xo.cpu().numpy(), c.layer0_x_out) < TOL
  assert rel_err(sk.cpu().numpy(), c.layer0_skip) < TOL


def test_layer_call_training_dropout_matches_reference_golden():
  """WaveNetLayer.call(training=True) (layers.py:178,192-196) with the fixture's injected keep-mask: forward against the
  reference-source golden, the adjoint (dx and every weight gradient) against the oracle with the same mask; training=False and a
  cleared mask behave as Keras does (no dropout / fresh masks per call)."""
  from oracle import wavenet_oracle as wo
  from wavenets_b200 import WaveNetLayer
  c = load_case('dropout_mask')
  kw = c.kw
  lay = WaveNetLayer(kernel=2, dilation_rate=[int(d) for d in c.dilations[0]], activation=kw['activation'], channels=kw['channels'],
                     skip_channels=kw['skip_channels'], dropout=kw['dropout'])
  x = c.layer0_x
  lay.build(x.shape)
  w = {k[len('block0/'):]: v for k, v in c.weights.items() if k.startswith('block0/')}
  lay.set_weights(w)
  lay.set_dropout_mask(c.keep_masks[0])
  xo, sk = lay.call(x, training=True)
  assert rel_err(xo.cpu().numpy(), c.layer0_x_out_train) < TOL
  assert rel_err(sk.cpu().numpy(), c.layer0_skip_train) < TOL
  # adjoint with seeded cotangents
  rng = np.random.default_rng(5)
  dxo, dsk = rng.standard_normal(xo.shape).astype(np.float32), rng.standard_normal(sk.shape).astype(np.float32)
  dx, _ = lay.backward(dxo, dsk)
  p = {k: v.astype(np.float64) for k, v in c.weights.items()}
  lc = wo._layer_cfgs(c.cfg)[0]
  _, _, cache = wo.layer_forward(p, 'block0', lc, x.astype(np.float64), None, keep=c.keep_masks[0], rate=c.cfg.dropout)
  dx_o, _, g_o = wo.layer_backward(p, 'block0', lc, cache, dxo.astype(np.float64), dsk.astype(np.float64))
  assert rel_err(dx.cpu().numpy(), dx_o) < TOL
  g = lay.get_grads()
  for k, ref in g_o.items():
    assert rel_err(g[k[len('block0/'):]], ref) < TOL, k
  # inference mode ignores the mask
  xo0, sk0 = lay.call(x, training=False)
  assert rel_err(xo0.cpu().numpy(), c.layer0_x_out) < TOL and rel_err(sk0.cpu().numpy(), c.layer0_skip) < TOL
  # built-in Philox masks: a new one per call, the conv branch really is masked
  lay.set_dropout_mask(None)
  a, _ = lay.call(x, training=True)
  b, _ = lay.call(x, training=True)
  assert float((a - b).abs().max()) > 1e-3
  assert float((a - xo0).abs().max()) > 1e-3


def test_saturated_softmax_takes_the_keras3_clip():
  """`cat_saturated` also runs in the parametrised test above; here: the clipped rows really are on the clipped branch (the
  per-row loss tops out at -log(1e-7) + log(sum of clipped probabilities)), through loss_fn on the model's own output too."""
  c = load_case('cat_saturated')
  m = _build(c, 'fp32')
  pred = m(c.x[:, :-1])
  per = m.loss_fn(m.prepare_target(c.x[:, 1:, :]), pred).cpu().numpy().reshape(c.B, c.T)
  assert rel_err(per, c.loss_per_sample) < TOL
  assert abs(per.max() + np.log(1e-7)) < 1e-3


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_dropout_philox_masks(precision):
  """Built-in masks: fresh per step (also under CUDA-graph replay), reproducible from the seed,
  inactive in test_step; bf16 tier takes the same path."""
  from wavenets_b200 import WaveNet
  kw = dict(channels=64, blocks=3, layers_per_block=1, dilation_bound=8, skip_channels=64, final_layers_channels=[64], dropout=0.3)
  rng = np.random.default_rng(3)
  x = np.clip(rng.standard_normal((2, 301, 1)) * 0.4, -1, 1).astype(np.float32)
  m = WaveNet(**kw, precision=precision)
  m.build(x[:, :-1].shape)
  m.handle.glorot_init(seed=1, bias_std=0.02)
  m.set_dropout_seed(11)
  a = [m.train_step(x)['loss'] for _ in range(4)]
  assert len(set(a)) == 4, a                       # a new mask every step (steps 3+ are graph replays)
  t0, t1 = m.test_step(x)['loss'], m.test_step(x)['loss']
  assert t0 == t1                                  # no dropout outside training
  m.set_dropout_seed(11)
  b = [m.train_step(x)['loss'] for _ in range(4)]
  assert a == b                                    # counter-based: same seed, same sequence
  m.set_dropout_seed(12)
  assert m.train_step(x)['loss'] != a[0]
  # injected all-ones mask == scaling the conv-branch input by 1/(1-p); all-zeros mask kills the conv branch input
  ones = [np.ones((2, 300, 64), bool)] * 3
  m.set_dropout_masks(ones)
  l1 = m.train_step(x)['loss']
  assert l1 == m.train_step(x)['loss']
  assert abs(l1 - t0) > 0
