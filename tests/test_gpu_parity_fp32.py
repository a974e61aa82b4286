"""GPU parity tests, fp32 tier: CUDA path (through the C ABI) vs the CPU oracle on the same
seeded inputs and weights.  Tolerance from the north star: <= 1e-4 relative in fp32;
bit-exact for the quantiser and causal-shift indexing."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as wo
from tests.util import COND_IN, SMALL_MODELS, make_inputs, oracle_config, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _model_and_oracle(kw, B, T, precision='fp32', seed=1):
  from wavenets_b200 import WaveNet
  cond_in = COND_IN if kw.get('conditioning') else 0
  cfg = oracle_config(kw, cond_in)
  p = wo.init_params(cfg, seed=seed)
  m = WaveNet(**kw, precision=precision)
  x, cond = make_inputs(B, T, cond_in)
  m.build((x[:, :-1].shape, cond.shape) if cond is not None else x[:, :-1].shape)
  assert m.variable_names == [n for n, _ in wo.param_specs(cfg)]
  m.set_weights({k: v.astype(np.float32) for k, v in p.items()})
  return m, cfg, p, x, cond


def test_quantizer_bit_exact():
  from wavenets_b200 import WaveNet
  m = WaveNet(channels=8, blocks=1, final_layers_channels=[], dilation_bound=2)
  rng = np.random.default_rng(0)
  edges = np.linspace(-1, 1, 257).astype(np.float32)
  x = np.concatenate([
    edges, np.nextafter(edges, np.float32(2)), np.nextafter(edges, np.float32(-2)),
    np.array([-1e-30, 1e-30, -0.0, 0.0, 1.0, -1.0, 0.99999994, -0.99999994], np.float32),
    rng.uniform(-1, 1, 100000).astype(np.float32)])
  got = m.prepare_target(x).cpu().numpy()
  assert got.dtype == np.int64
  assert np.array_equal(got, wo.discretize(x, 8))
  m16 = WaveNet(channels=8, blocks=1, final_layers_channels=[], dilation_bound=2, bits=16)
  e16 = np.linspace(-1, 1, 2 ** 16 + 1).astype(np.float32)
  x16 = np.concatenate([e16, np.nextafter(e16, np.float32(2)), np.nextafter(e16, np.float32(-2)), x])
  assert np.array_equal(m16.prepare_target(x16).cpu().numpy(), wo.discretize(x16, 16))
  # empty input
  assert m.prepare_target(np.zeros((0,), np.float32)).numel() == 0


@pytest.mark.parametrize('name', sorted(SMALL_MODELS))
@pytest.mark.parametrize('BT', [(2, 24), (3, 151)])
def test_train_step_matches_oracle(name, BT):
  kw = SMALL_MODELS[name]
  B, T = BT
  m, cfg, p, x, cond = _model_and_oracle(kw, B, T)
  loss_o, g_o, aux = wo.train_step(p, cfg, x.astype(np.float64), None if cond is None else cond.astype(np.float64))
  out = m.train_step((x, cond) if cond is not None else x)
  # the reference reports the regulariser separately (metrics 'loss' and 'reg_loss', model.py:340-344)
  assert abs(out['loss'] - aux['loss_no_reg']) <= TOL * abs(loss_o), (out['loss'], aux['loss_no_reg'])
  assert abs(out.get('reg_loss', 0.0) - aux['reg_loss']) <= TOL * abs(loss_o)
  assert ('reg_loss' in out) == (cfg.l2_reg_factor > 0)
  g = m.get_grads()
  for k in g_o:
    assert rel_err(g[k], g_o[k]) < TOL, (k, rel_err(g[k], g_o[k]))
  # forward output (probabilities / mixture params)
  pred_o, _ = wo.model_forward(p, cfg, x[:, :-1].astype(np.float64), None if cond is None else cond.astype(np.float64))
  pred = m((x[:, :-1], cond) if cond is not None else x[:, :-1]).cpu().numpy()
  assert pred.shape == pred_o.shape
  assert rel_err(pred, pred_o) < TOL
  # test_step = same loss without gradients
  assert abs(m.test_step((x, cond) if cond is not None else x)['loss'] - aux['loss_no_reg']) <= TOL * abs(loss_o)


@pytest.mark.parametrize('bits,scale', [(6, 150.0), (8, 300.0)])
def test_saturated_softmax_keras3_clip(bits, scale):
  """Keras 3's clip inside sparse_categorical_crossentropy (model.py:516) on a saturated softmax — logits conv scaled until a
  large part of the class AND target probabilities leave [1e-7, 1 - 1e-7] — against the oracle (itself pinned to the reference-source
  golden `cat_saturated`): 64 classes run the generic loss kernel, 256 the register-resident one."""
  from wavenets_b200 import WaveNet
  kw = dict(channels=8, blocks=2, layers_per_block=1, dilation_bound=4, final_layers_channels=[16, 16], activation='tanh', bits=bits)
  B, T = 2, 120
  cfg = oracle_config(kw, 0)
  p = {k: v.astype(np.float32).astype(np.float64) for k, v in wo.init_params(cfg, seed=1).items()}
  p['final2/kernel'] = (p['final2/kernel'] * scale).astype(np.float32).astype(np.float64)
  x, _ = make_inputs(B, T, 0)
  m = WaveNet(**kw)
  m.build(x[:, :-1].shape)
  m.set_weights({k: v.astype(np.float32) for k, v in p.items()})
  loss_o, g_o, aux = wo.train_step(p, cfg, x.astype(np.float64), None)
  pred_o, _ = wo.model_forward(p, cfg, x[:, :-1].astype(np.float64), None)
  tgt = wo.discretize(x[:, 1:, 0], bits)
  p_y = np.take_along_axis(pred_o, tgt[..., None], -1)[..., 0]
  assert 0.1 < (p_y < 1e-7).mean() < 0.9 and (pred_o < 1e-7).mean() > 0.1          # clipped and un-clipped targets both present
  assert abs(aux['loss_per_sample'].max() + np.log(1e-7)) < 1e-3                    # the clipped rows sit at -log(1e-7) (+ log sum)
  out = m.train_step(x)
  assert abs(out['loss'] - aux['loss_no_reg']) <= TOL * abs(loss_o), (out['loss'], aux['loss_no_reg'])
  g = m.get_grads()
  for k in g_o:
    assert rel_err(g[k], g_o[k]) < TOL, (k, rel_err(g[k], g_o[k]))
  assert abs(m.test_step(x)['loss'] - aux['loss_no_reg']) <= TOL * abs(loss_o)


LAYERS = {
  'single': dict(dilation_rate=4, channels=8),
  'multi_leaky': dict(dilation_rate=[1, 2, 4], activation='leaky_relu', channels=8, dilation_channels=12, skip_channels=6),
  'cond': dict(dilation_rate=2, channels=8, condition=True, skip_channels=4),
  'nores': dict(dilation_rate=[3, 1], activation='tanh', channels=8, residual=False),
  'k3': dict(kernel=3, dilation_rate=[1, 3], activation='relu', channels=8),
}


@pytest.mark.parametrize('name', sorted(LAYERS))
def test_layer_call_and_adjoint_match_oracle(name):
  from wavenets_b200 import WaveNetLayer
  kw = LAYERS[name]
  B, T, R = 2, 37, kw['channels']
  rng = np.random.default_rng(3)
  lay = WaveNetLayer(**kw)
  x = rng.standard_normal((B, T, R)).astype(np.float32)
  Cc = 5
  cond = rng.standard_normal((B, Cc)).astype(np.float32) if kw.get('condition') else None
  cond3 = None if cond is None else np.repeat(cond[:, None, :], T, axis=1)
  lay.build((x.shape, cond3.shape) if cond is not None else x.shape)
  dils = kw['dilation_rate'] if isinstance(kw['dilation_rate'], list) else [kw['dilation_rate']]
  D = kw.get('dilation_channels') or R
  S = kw.get('skip_channels')
  K = kw.get('kernel', 2)
  # random weights in Keras layouts
  w = {}
  for n, shape in zip(lay.weight_names, lay._handle.shapes):
    w[n] = (rng.standard_normal(shape) * 0.3).astype(np.float32)
  lay.set_weights(w)
  assert lay.compute_output_shape(x.shape) == ((B, T, R), (B, T, S or R))
  xo, sk = lay((x, cond3) if cond is not None else x)
  p = {'block0/' + k: v.astype(np.float64) for k, v in w.items()}
  lc = dict(dilations=dils, activation=kw.get('activation'), residual=kw.get('residual', True), has_skip=S is not None,
            condition=cond is not None)
  xo_o, sk_o, cache = wo.layer_forward(p, 'block0', lc, x.astype(np.float64), None if cond is None else cond.astype(np.float64))
  assert rel_err(xo.cpu().numpy(), xo_o) < TOL
  assert rel_err(sk.cpu().numpy(), sk_o) < TOL
  dxo = rng.standard_normal(xo_o.shape).astype(np.float32)
  dsk = rng.standard_normal(sk_o.shape).astype(np.float32)
  dx, dcond = lay.backward(dxo, dsk)
  dx_o, dcond_o, g_o = wo.layer_backward(p, 'block0', lc, cache, dxo.astype(np.float64), dsk.astype(np.float64))
  assert rel_err(dx.cpu().numpy(), dx_o) < TOL
  if cond is not None:
    assert rel_err(dcond.cpu().numpy(), dcond_o) < TOL
  g = lay.get_grads()
  for k, v in g_o.items():
    assert rel_err(g[k[len('block0/'):]], v) < TOL, k


def test_layer_errors_match_reference():
  from wavenets_b200 import WaveNetLayer
  lay = WaveNetLayer(dilation_rate=1, channels=8, condition=True)
  with pytest.raises(ValueError, match='same length'):
    lay.build(((1, 16, 8), (1, 15, 4)))
  lay2 = WaveNetLayer(dilation_rate=1, channels=8)
  with pytest.raises(ValueError, match='Residual connection'):
    lay2.build((1, 16, 12))


def test_causality_receptive_field_and_batch_isolation():
  kw = dict(channels=8, blocks=3, layers_per_block=1, dilation_bound=4, final_layers_channels=[8])
  m, cfg, p, x, _ = _model_and_oracle(kw, 2, 40)
  rf = m.receptive_field
  xin = x[:, :-1].copy()
  y0 = m(xin).cpu().numpy()
  t = 30
  for dt, changed in [(1, False), (0, True), (-(rf - 1), True), (-rf, False)]:
    x2 = xin.copy()
    x2[0, t + dt, 0] += 0.5
    y2 = m(x2).cpu().numpy()
    assert (np.abs(y2[0, t] - y0[0, t]).max() > 0) == changed, (dt, changed)
    assert np.array_equal(y2[1], y0[1])     # the other batch row is bit-identical
  # causal shift indexing is bit-exact: the first T-1 outputs do not depend on the last input sample
  x3 = xin.copy()
  x3[:, -1, 0] = 0.7
  y3 = m(x3).cpu().numpy()
  assert np.array_equal(y3[:, :-1], y0[:, :-1])


def test_defaults_topology_c1_small_segment():
  # configs[0]: defaults.yaml topology, 256-way softmax, batch 1 (short segment so the oracle runs in seconds)
  from wavenets_b200 import CONFIGS, model_kwargs
  kw = model_kwargs(CONFIGS['c1'])
  m, cfg, p, x, _ = _model_and_oracle(kw, 1, 1000)
  assert m.receptive_field == 768 and m.handle.n_scalars >= 170816
  loss_o, g_o, _ = wo.train_step(p, cfg, x.astype(np.float64))
  out = m.train_step(x)
  assert abs(out['loss'] - loss_o) <= TOL * abs(loss_o)
  g = m.get_grads()
  for k in g_o:
    assert rel_err(g[k], g_o[k]) < TOL, k


def test_run_to_run_determinism():
  kw = SMALL_MODELS['cond_skip']
  m, cfg, p, x, cond = _model_and_oracle(kw, 3, 200)
  m.train_step((x, cond))
  g1 = m.handle.flat_grads.clone()
  m.train_step((x, cond))
  assert torch.equal(g1, m.handle.flat_grads)   # no atomics anywhere: bit-stable gradients
