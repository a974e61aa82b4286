"""autoregressive generation with per-conv input histories (model.py:258-307, layers.py:226-290)."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as wo
from tests.util import COND_IN, SMALL_MODELS, make_inputs, oracle_config, rel_err

pytestmark = pytest.mark.gpu


def _model(name, precision='fp32'):
  from wavenets_b200 import WaveNet
  kw = SMALL_MODELS[name]
  cond_in = COND_IN if kw.get('conditioning') else 0
  cfg = oracle_config(kw, cond_in)
  p = wo.init_params(cfg, seed=1)
  m = WaveNet(**kw, precision=precision)
  x, cond = make_inputs(3, 90, cond_in)
  m.build((x[:, :-1].shape, cond.shape) if cond is not None else x[:, :-1].shape)
  m.set_weights({k: v.astype(np.float32) for k, v in p.items()})
  return m, cfg, p, x, cond


@pytest.mark.parametrize('name', sorted(SMALL_MODELS))
def test_step_form_equals_full_forward(name):
  """Teacher-forced, the single-step form reproduces WaveNet.call position by position (and the oracle)."""
  m, cfg, p, x, cond = _model(name)
  n_prime = 17
  pred = m._teacher_forced_step_predictions(x[:, :n_prime], x[:, n_prime:-1], cond).cpu().numpy()
  xin = x[:, :-1]
  full = m((xin, cond) if cond is not None else xin).cpu().numpy()
  # the prediction made at position t (inputs up to t) is row t of the full forward; generated position n_prime + i uses t = n_prime - 1 + i
  assert rel_err(pred, full[:, n_prime - 1:-1]) < 1e-4
  ref, _ = wo.model_forward({k: v.astype(np.float64) for k, v in p.items()}, cfg, xin.astype(np.float64), None if cond is None else cond.astype(np.float64))
  assert rel_err(pred, ref[:, n_prime - 1:-1]) < 1e-4


@pytest.mark.parametrize('name', ['logistic_cond', 'gaussian_noskip', 'categorical_multidil'])
def test_generate_matches_windowed_reference_loop(name):
  """Deterministic generation == the reference's loop: forward over the sliding receptive-field window, take the last
  position, sample deterministically, append (model.py:296-305)."""
  m, cfg, p, x, cond = _model(name)
  rf = m.receptive_field
  B, L = 2, 12
  prime = x[:B, :rf]
  c = None if cond is None else cond[:B]
  got = m.generate(L, condition=c, sample=prime, deterministic=True).cpu().numpy()
  assert got.shape == (B, L, 1)
  p64 = {k: v.astype(np.float64) for k, v in p.items()}
  win = prime.astype(np.float64)
  outs = []
  for _ in range(L):
    pred, _ = wo.model_forward(p64, cfg, win, None if c is None else c.astype(np.float64))
    s = wo.sample_deterministic(cfg, pred[:, -1:, :]).reshape(B, 1, 1)
    outs.append(s)
    win = np.concatenate([win[:, 1:], s], axis=1)
  ref = np.concatenate(outs, axis=1)
  if cfg.num_mixtures is None:
    assert (np.abs(got - ref) < 1e-6).mean() >= 0.9       # an argmax near-tie may send one trajectory elsewhere
  else:
    assert np.abs(got - ref).max() < 1e-3


def test_generate_stochastic_and_errors():
  m, cfg, p, x, cond = _model('cond_skip')
  with pytest.raises(ValueError, match='Conditioning must be provided'):
    m.generate(4)
  a = m.generate(16, condition=cond, seed=5).cpu().numpy()
  b = m.generate(16, condition=cond, seed=5).cpu().numpy()
  c = m.generate(16, condition=cond, seed=6).cpu().numpy()
  assert a.shape == (3, 16, 1) and np.array_equal(a, b) and not np.array_equal(a, c)
  assert np.abs(a).max() <= 1.0
  with pytest.raises(ValueError, match='same batch size'):
    m.generate(4, condition=cond, sample=x[:2, :10])
