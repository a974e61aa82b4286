"""The CPU oracle against the golden vectors produced by the REFERENCE'S OWN SOURCE
(oracle/make_golden.py: /root/reference/src/{layers,model}.py executed unmodified over
oracle/tf_shim, float64).  Nothing here reads /root/reference: the fixtures are committed."""
import glob
import os

import numpy as np
import pytest

from oracle import wavenet_oracle as wo
from tests.golden_util import CASES, load_case

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_fixtures_present():
  names = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, '*.npz'))}
  assert set(CASES) | {'known_answers'} <= names


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_matches_reference_golden(name):
  c = load_case(name)
  cfg, p = c.cfg, {k: v.astype(np.float64) for k, v in c.weights.items()}
  x = c.x.astype(np.float64)
  cond = None if c.cond is None else c.cond.astype(np.float64)
  # structure
  per_block, rf = wo.dilation_schedule(cfg)
  assert rf == c.receptive_field
  assert np.array_equal(np.array(per_block), c.dilations)
  assert [n for n, _ in wo.param_specs(cfg)] == list(c.weights)          # Keras variable order
  # forward (inference mode)
  pred, _ = wo.model_forward(p, cfg, x[:, :-1], cond)
  np.testing.assert_allclose(pred[:, c.pred_t], c.pred, rtol=2e-6, atol=1e-9)
  # targets: bit-exact
  if cfg.num_mixtures is None:
    assert np.array_equal(wo.discretize(x[:, 1:, 0], cfg.bits), c.target[..., 0])
  # deterministic sampling branch
  np.testing.assert_allclose(wo.sample_deterministic(cfg, pred), c.sample_deterministic, rtol=1e-6, atol=1e-7)
  # test_step: loss without dropout
  loss_t, _, aux_t = wo.train_step(p, cfg, x, cond, n_replicas=c.n_replicas)
  assert abs(aux_t['loss_no_reg'] - c.test_loss) <= 1e-9 * abs(c.test_loss)
  # per-(b,t) losses: the reference's categorical path clips probabilities to [1e-7, 1-1e-7] and renormalises (Keras 3
  # sparse_categorical_crossentropy); the oracle restates exactly that (case `cat_saturated` has the clip active)
  np.testing.assert_allclose(aux_t['loss_per_sample'], c.loss_per_sample, rtol=1e-9, atol=1e-9)
  # train_step: loss parts and every gradient that reaches optimizer.apply_gradients
  loss, g, aux = wo.train_step(p, cfg, x, cond, n_replicas=c.n_replicas, keep_masks=c.keep_masks)
  assert abs(aux['loss_no_reg'] - c.train_loss) <= 1e-9 * abs(c.train_loss)
  if c.reg_loss is not None:
    assert abs(aux['reg_loss'] - c.reg_loss) <= 1e-6 * abs(c.reg_loss)      # fixture weights are fp32-rounded
  assert set(g) == set(c.grads)
  for k in c.grads:
    scale = np.abs(c.grads[k]).max() + 1e-30
    err = np.abs(g[k] - c.grads[k]).max() / scale
    assert err < 2e-6, (k, err)      # fixtures store fp32


@pytest.mark.parametrize('name', sorted(CASES))
def test_layer_call_matches_reference_golden(name):
  c = load_case(name)
  cfg, p = c.cfg, {k: v.astype(np.float64) for k, v in c.weights.items()}
  lc = wo._layer_cfgs(cfg)[0]
  cond = None if c.layer0_cond is None else c.layer0_cond.astype(np.float64)
  x_out, skip, _ = wo.layer_forward(p, 'block0', lc, c.layer0_x.astype(np.float64), cond)
  np.testing.assert_allclose(x_out, c.layer0_x_out, rtol=2e-6, atol=2e-7)
  np.testing.assert_allclose(skip, c.layer0_skip, rtol=2e-6, atol=2e-7)
  if c.layer0_x_out_train is not None:
    # WaveNetLayer.call(training=True) with the injected keep-mask of block 0 (layers.py:192-196)
    x_out, skip, _ = wo.layer_forward(p, 'block0', lc, c.layer0_x.astype(np.float64), cond, keep=c.keep_masks[0], rate=cfg.dropout)
    np.testing.assert_allclose(x_out, c.layer0_x_out_train, rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(skip, c.layer0_skip_train, rtol=2e-6, atol=2e-7)
    assert np.abs(c.layer0_x_out_train - c.layer0_x_out).max() > 1e-3


def test_saturated_case_has_the_clip_active():
  c = load_case('cat_saturated')
  assert c.clip_stats[0] > 0.2 and 0.2 < c.clip_stats[1] < 0.8     # clipped probabilities / clipped TARGET probabilities
  # rows whose target probability is clipped carry exactly -log(1e-7) + log(sum of clipped probabilities)
  assert np.isclose(c.loss_per_sample.max(), -np.log(1e-7), atol=1e-4)


def test_known_answers():
  ka = np.load(os.path.join(GOLD, 'known_answers.npz'))
  cfg = wo.Config(kernel_size=2, channels=32, blocks=5, layers_per_block=5, dilation_bound=256, activation='leaky_relu',
                  final_layers_channels=[128, 256])
  per_block, rf = wo.dilation_schedule(cfg)
  assert rf == int(ka['defaults_rf']) == 768
  assert np.array_equal(np.array(per_block), ka['defaults_dilations'])
  assert wo.dilation_schedule(wo.Config(blocks=40, layers_per_block=1, dilation_bound=1024, use_skip=False))[1] == int(ka['c5_rf'])
  assert wo.dilation_schedule(wo.Config(blocks=30, layers_per_block=1, dilation_bound=1024))[1] == int(ka['c2_rf'])
  for bits in (8, 16):
    assert np.array_equal(wo.discretize(ka[f'quant{bits}_x'], bits), ka[f'quant{bits}_idx'])


def test_constructor_errors_match_reference():
  """Same ValueError messages as model.py:52-70, from the host mirror (no GPU needed to construct)."""
  from wavenets_b200 import WaveNet
  ka = np.load(os.path.join(GOLD, 'known_answers.npz'))
  bad = [dict(conditioning='x'), dict(kernel_size=1), dict(dilation_bound=100), dict(layers_per_block=0), dict(blocks=0),
         dict(num_mixtures=0, sampling_function='logistic'), dict(dropout=1.5), dict(sampling_function='foo'),
         dict(sampling_function='categorical', num_mixtures=3)]
  for kw, msg in zip(bad, ka['value_errors']):
    assert str(msg) != ''
    with pytest.raises(ValueError) as e:
      WaveNet(**{**dict(final_layers_channels=[]), **kw})
    assert str(e.value).replace(' ', '') == str(msg).replace(' ', '')
    with pytest.raises(ValueError):
      wo.Config(**{**dict(final_layers_channels=[]), **kw}).validate()
