"""GPU parity tests, bf16 tier (tcgen05/TMEM/TMA): CUDA path through the C ABI vs the fp64 CPU
oracle on the same seeded inputs and weights.

Stated bf16 tolerance (activations and GEMM operands are stored in bf16, accumulation is fp32,
tanh/sigmoid use MUFU.TANH), against the fp64 oracle:
  * outputs: relative L2 <= 1e-2; loss: <= 1e-2 relative;
  * gradients, smooth activations (None/tanh/sigmoid): relative L2 <= 1.5e-2 per tensor
    (measured 3e-3 .. 6e-3);
  * gradients, piecewise-linear activations (relu/leaky_relu): relative L2 <= 1e-1 per tensor AND
    cosine similarity >= 0.995.  The derivative of relu/leaky_relu is discontinuous: the ~0.25 %
    of pre-activations that lie within one bf16 ulp of zero change sign under bf16 rounding,
    each flipping a derivative by 0.8..1.0, which alone is a relative-L2 error of
    sqrt(0.0025)*0.8 ~ 4 % — a property of bf16 storage, not of the kernels (the same models
    with a smooth activation meet the tight bound, see `multidil_alias_tanh`).
The gate that a kernel bug cannot hide behind is the bf16-FAITHFUL oracle (oracle/faithful.py: the same model rounding to
bf16 exactly where the kernels do, relu / leaky_relu derivative masks read back from the device — tests/util.py:
device_slope_masks): every gradient tensor within TOL_GRAD_FAITHFUL relative L2 for EVERY activation."""
import numpy as np
import pytest
import torch

from oracle import faithful
from oracle import wavenet_oracle as wo
from tests.util import device_slope_masks, make_inputs, oracle_config, rel_l2

pytestmark = pytest.mark.gpu
TOL_OUT, TOL_LOSS, TOL_GRAD, TOL_GRAD_PWL, MIN_COS_PWL = 1e-2, 1e-2, 1.5e-2, 1e-1, 0.995
TOL_GRAD_FAITHFUL, TOL_LOSS_FAITHFUL = 8e-3, 3e-4   # measured <= 5.6e-3 (causal/kernel of a 200-row batch)
COND_IN = 9

BF16_MODELS = {
  'single_dil_skip_cond': dict(channels=64, blocks=4, layers_per_block=1, dilation_bound=8, skip_channels=64,
                               final_layers_channels=[64], conditioning='global', mapping_layers=[8, 16],
                               mapping_activation='leaky_relu', activation='leaky_relu'),
  'multidil_alias': dict(channels=64, blocks=2, layers_per_block=3, dilation_bound=8, activation='leaky_relu',
                         final_layers_channels=[128, 64]),
  'multidil_alias_tanh': dict(channels=64, blocks=2, layers_per_block=3, dilation_bound=8, activation='tanh',
                              final_layers_channels=[128, 64]),
  'wide_noskip': dict(channels=128, blocks=3, layers_per_block=1, dilation_bound=4, use_skip=False,
                      final_layers_channels=[128], dilation_channels=64),
  'logistic': dict(channels=64, blocks=3, layers_per_block=1, dilation_bound=4, skip_channels=128,
                   final_layers_channels=[64], num_mixtures=10, sampling_function='logistic', bits=16,
                   conditioning='global', mapping_layers=[8], mapping_activation='relu', activation='relu'),
  'gaussian_k3': dict(channels=64, blocks=2, layers_per_block=2, dilation_bound=9, kernel_size=3, activation='tanh',
                      final_layers_channels=[64], num_mixtures=4, sampling_function='gaussian'),
  'r256': dict(channels=256, blocks=2, layers_per_block=1, dilation_bound=4, skip_channels=256,
               final_layers_channels=[256], activation='leaky_relu'),
  # shapes the fused block-forward kernel takes (gemm_tc_block.cuh): D = R in {128, 256}, residual on
  'fused128_multidil_cond': dict(channels=128, blocks=2, layers_per_block=2, dilation_bound=8, skip_channels=128, activation='tanh',
                                 final_layers_channels=[128], conditioning='global', mapping_layers=[8], mapping_activation='tanh'),
  'fused256_k3_alias': dict(channels=256, blocks=2, layers_per_block=1, dilation_bound=9, kernel_size=3, final_layers_channels=[128]),
  # shapes whose block weight gradients take the grouped launch (gemm_tc_wgroup.cuh): R, D, S multiples of 256
  'group256_cond_l2': dict(channels=256, blocks=3, layers_per_block=1, dilation_bound=4, skip_channels=256, final_layers_channels=[256],
                           activation='tanh', conditioning='global', mapping_layers=[8], mapping_activation='tanh', l2_reg_factor=0.01),
  'group256_noskip_k3': dict(channels=256, blocks=2, layers_per_block=1, dilation_bound=9, kernel_size=3, use_skip=False,
                             final_layers_channels=[128]),
  'group256_multidil': dict(channels=256, blocks=2, layers_per_block=2, dilation_bound=8, skip_channels=256, final_layers_channels=[128],
                            activation='tanh'),
  'group256_alias_multidil': dict(channels=256, blocks=2, layers_per_block=3, dilation_bound=8, final_layers_channels=[128], activation='tanh'),
  # the reference's default width (defaults.yaml:10): 32 channels — k-blocks padded to 64 in shared memory by TMA zero fill
  'narrow32_defaults': dict(channels=32, blocks=2, layers_per_block=3, dilation_bound=8, activation='leaky_relu', final_layers_channels=[128, 256]),
  'narrow32_cond_skip96': dict(channels=32, blocks=3, layers_per_block=1, dilation_bound=8, skip_channels=96, dilation_channels=64,
                               final_layers_channels=[32], conditioning='global', mapping_layers=[8], mapping_activation='tanh', activation='tanh'),
  'narrow32_k3_gaussian': dict(channels=32, blocks=2, layers_per_block=2, dilation_bound=9, kernel_size=3, activation='tanh',
                               final_layers_channels=[96], num_mixtures=4, sampling_function='gaussian', use_skip=False),
  'unfused256_nores': dict(channels=256, blocks=2, layers_per_block=1, dilation_bound=4, skip_channels=128, use_residual=False,
                           final_layers_channels=[128]),
}


def _smooth(kw):
  return all(kw.get(k) in (None, 'tanh', 'sigmoid') for k in ('activation', 'mapping_activation'))


def _cos(a, b):
  a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
  return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


def _build(kw, B, T):
  from wavenets_b200 import WaveNet
  cond_in = COND_IN if kw.get('conditioning') else 0
  cfg = oracle_config(kw, cond_in)
  p = wo.init_params(cfg, seed=1)
  m = WaveNet(**kw, precision='bf16')
  x, cond = make_inputs(B, T, cond_in)
  m.build((x[:, :-1].shape, cond.shape) if cond is not None else x[:, :-1].shape)
  m.set_weights({k: v.astype(np.float32) for k, v in p.items()})
  return m, cfg, p, x, cond


@pytest.mark.parametrize('name', sorted(BF16_MODELS))
@pytest.mark.parametrize('BT', [(2, 100), (3, 333)])
def test_train_step_matches_oracle_bf16(name, BT):
  kw = BF16_MODELS[name]
  B, T = BT
  m, cfg, p, x, cond = _build(kw, B, T)
  c64 = None if cond is None else cond.astype(np.float64)
  loss_o, g_o, aux = wo.train_step(p, cfg, x.astype(np.float64), c64)
  data = (x, cond) if cond is not None else x
  out = m.train_step(data)
  # the reference reports the regulariser separately (metrics 'loss' and 'reg_loss', model.py:340-344)
  assert abs(out['loss'] - aux['loss_no_reg']) <= TOL_LOSS * abs(loss_o), (out['loss'], loss_o)
  assert abs(out.get('reg_loss', 0.0) - aux['reg_loss']) <= TOL_LOSS * abs(loss_o)
  g = m.get_grads()
  worst = max((rel_l2(g[k], g_o[k]), k) for k in g_o if np.linalg.norm(g_o[k]) > 0)
  if _smooth(kw):
    assert worst[0] < TOL_GRAD, worst
  else:
    assert worst[0] < TOL_GRAD_PWL, worst
    cos = min((_cos(g[k], g_o[k]), k) for k in g_o if np.linalg.norm(g_o[k]) > 0)
    assert cos[0] > MIN_COS_PWL, cos
  # bf16-faithful oracle: same roundings as the kernels -> tight for every activation
  p32 = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
  l_f, g_f = faithful.train_step(p32, cfg, x, cond, faithful=True, slope_masks=device_slope_masks(m, kw, B, T))
  assert abs(out['loss'] - l_f) <= TOL_LOSS_FAITHFUL * abs(l_f), (out['loss'], l_f)
  scale = max(np.linalg.norm(v) for v in g_f.values())
  worst_f = max((rel_l2(g[k], g_f[k]), k) for k in g_f if np.linalg.norm(g_f[k]) > 1e-9 * scale)
  assert worst_f[0] < TOL_GRAD_FAITHFUL, worst_f
  pred_o, _ = wo.model_forward(p, cfg, x[:, :-1].astype(np.float64), c64)
  pred = m((x[:, :-1], cond) if cond is not None else x[:, :-1]).cpu().numpy()
  assert rel_l2(pred, pred_o) < TOL_OUT
  assert rel_l2(pred, faithful.forward(p32, cfg, x[:, :-1], cond, faithful=True)) < 3e-3
  assert abs(m.test_step(data)['loss'] - aux['loss_no_reg']) <= TOL_LOSS * abs(loss_o)


def test_bf16_layer_call_and_adjoint():
  from wavenets_b200 import WaveNetLayer
  rng = np.random.default_rng(5)
  B, T, R = 2, 200, 64
  lay = WaveNetLayer(dilation_rate=[1, 4], activation='leaky_relu', channels=R, skip_channels=128, precision='bf16')
  x = rng.standard_normal((B, T, R)).astype(np.float32)
  lay.build(x.shape)
  w = {n: (rng.standard_normal(s) * 0.1).astype(np.float32) for n, s in zip(lay.weight_names, lay._handle.shapes)}
  lay.set_weights(w)
  xo, sk = lay(x)
  p = {'block0/' + k: v.astype(np.float64) for k, v in w.items()}
  lc = dict(dilations=[1, 4], activation='leaky_relu', residual=True, has_skip=True, condition=False)
  xo_o, sk_o, cache = wo.layer_forward(p, 'block0', lc, x.astype(np.float64), None)
  assert rel_l2(xo.cpu().numpy(), xo_o) < TOL_OUT
  assert rel_l2(sk.cpu().numpy(), sk_o) < TOL_OUT
  dxo = rng.standard_normal(xo_o.shape).astype(np.float32)
  dsk = rng.standard_normal(sk_o.shape).astype(np.float32)
  dx, _ = lay.backward(dxo, dsk)
  dx_o, _, g_o = wo.layer_backward(p, 'block0', lc, cache, dxo.astype(np.float64), dsk.astype(np.float64))
  assert rel_l2(dx.cpu().numpy(), dx_o) < TOL_GRAD_PWL
  g = lay.get_grads()
  for k, v in g_o.items():
    assert rel_l2(g[k[len('block0/'):]], v) < TOL_GRAD_PWL, k
    assert _cos(g[k[len('block0/'):]], v) > MIN_COS_PWL, k


def test_bf16_causality_and_batch_isolation_exact():
  kw = dict(channels=64, blocks=3, layers_per_block=1, dilation_bound=4, final_layers_channels=[64])
  m, cfg, p, x, _ = _build(kw, 2, 300)
  rf = m.receptive_field
  xin = x[:, :-1].copy()
  y0 = m(xin).cpu().numpy()
  t = 200
  for dt, changed in [(1, False), (0, True), (-(rf - 1), True), (-rf, False)]:
    x2 = xin.copy()
    x2[0, t + dt, 0] += 0.5
    y2 = m(x2).cpu().numpy()
    assert (np.abs(y2[0, t] - y0[0, t]).max() > 0) == changed, (dt, changed)
    assert np.array_equal(y2[1], y0[1])


@pytest.mark.parametrize('name', ['single_dil_skip_cond', 'r256', 'fused128_multidil_cond'])
def test_bf16_determinism(name):
  """Bit-identical gradients run to run (no atomics anywhere; also through CUDA-graph replays, steps 3+)."""
  kw = BF16_MODELS[name]
  m, cfg, p, x, cond = _build(kw, 3, 700)
  data = (x, cond) if cond is not None else x
  m.train_step(data)
  g1 = m.handle.flat_grads.clone()
  for _ in range(3):
    m.train_step(data)
    assert torch.equal(g1, m.handle.flat_grads)


def test_bf16_rejects_unaligned_widths():
  from wavenets_b200 import WaveNet
  m = WaveNet(channels=40, blocks=2, final_layers_channels=[64], dilation_bound=4, precision='bf16')
  with pytest.raises(NotImplementedError):
    m.build((1, 64))
