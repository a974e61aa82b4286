"""WaveNet.sample_waveform (model.py:393-503) and the compiled MSE metric (model.py:338-346).
Deterministic branch: against the golden vectors from the reference's own source.  Stochastic branch: TF's
stateless seed-(4,2) stream cannot be reproduced, so parity is statistical (distribution of the draws)."""
import numpy as np
import pytest
import torch

from tests.golden_util import load_case

pytestmark = pytest.mark.gpu


def _build(c, precision='fp32'):
  from wavenets_b200 import WaveNet
  m = WaveNet(**c.kw, precision=precision)
  x_in = c.x[:, :-1]
  m.build((x_in.shape, c.cond.shape) if c.cond is not None else x_in.shape)
  m.set_weights(c.weights)
  return m


@pytest.mark.parametrize('name', ['cat_multidil', 'cond_skip', 'logistic_cond', 'gaussian_noskip_k3'])
def test_deterministic_sampling_matches_reference_golden(name):
  c = load_case(name)
  m = _build(c)
  x_in = c.x[:, :-1]
  pred = m((x_in, c.cond) if c.cond is not None else x_in)
  got = m.sample_waveform(pred, deterministic=True).cpu().numpy()
  ref = c.sample_deterministic
  assert got.shape == ref.shape
  close = np.abs(got - ref) <= 1e-4
  assert close.mean() > 0.995, close.mean()      # an fp32-vs-fp64 near-tie may pick the neighbouring bin / component


def test_categorical_draws_follow_the_distribution():
  from wavenets_b200 import WaveNet
  m = WaveNet(channels=8, blocks=1, final_layers_channels=[], dilation_bound=2, bits=8)
  B, T = 64, 4096
  m.build((B, T, 1))
  rng = np.random.default_rng(0)
  p = rng.dirichlet(np.full(256, 0.3)).astype(np.float32)
  pred = torch.from_numpy(np.broadcast_to(p, (B, T, 256)).copy()).cuda()
  s = m.sample_waveform(pred, seed=7)
  assert s.shape == (B, T, 1)
  idx = torch.round((s[..., 0] + 1.0) * 128.0).long().flatten().cpu().numpy()
  assert idx.min() >= 0 and idx.max() <= 255
  n = idx.size
  freq = np.bincount(idx, minlength=256) / n
  sigma = np.sqrt(p * (1 - p) / n) + 1e-9
  assert np.abs(freq - p).max() < 6 * sigma.max() and (np.abs(freq - p) / sigma).max() < 7
  # a different seed gives different draws; the same seed reproduces them
  assert not torch.equal(s, m.sample_waveform(pred, seed=8))
  assert torch.equal(s, m.sample_waveform(pred, seed=7))


@pytest.mark.parametrize('fn', ['gaussian', 'logistic'])
def test_mixture_draws_follow_the_distribution(fn):
  from wavenets_b200 import WaveNet
  M = 3
  m = WaveNet(channels=8, blocks=1, final_layers_channels=[], dilation_bound=2, num_mixtures=M, sampling_function=fn, bits=16)
  B, T = 32, 8192
  m.build((B, T, 1))
  w = np.log(np.array([0.2, 0.5, 0.3], np.float32))
  mu = np.array([-0.4, 0.1, 0.5], np.float32)
  # logistic tails are heavy: keep the components > 30 scale units apart so nearest-mean assignment is clean
  ls = np.log(np.array([0.02, 0.03, 0.01], np.float32) * (1.0 if fn == 'gaussian' else 0.2))
  pred = torch.from_numpy(np.broadcast_to(np.concatenate([w, mu, ls]), (B, T, 3 * M)).copy()).cuda()
  s = m.sample_waveform(pred, seed=3)[..., 0].flatten().cpu().numpy()
  comp = np.argmin(np.abs(s[:, None] - mu[None, :]), axis=1)       # components are > 10 sigma apart
  frac = np.bincount(comp, minlength=M) / s.size
  assert np.abs(frac - np.array([0.2, 0.5, 0.3])).max() < 0.01
  for k in range(M):
    sk = s[comp == k]
    scale = float(np.exp(ls[k]))
    std = scale if fn == 'gaussian' else scale * np.pi / np.sqrt(3.0)
    assert abs(sk.mean() - mu[k]) < 5 * std / np.sqrt(sk.size) + 1e-4
    assert abs(sk.std() - std) < 0.05 * std
  det = m.sample_waveform(pred, deterministic=True)
  assert det.shape == (B, T, 1) and torch.allclose(det, torch.full_like(det, 0.1))


def test_mse_metric_inside_train_step():
  from wavenets_b200 import WaveNet
  from wavenets_b200.metrics import MeanSquaredError
  c = load_case('cat_multidil')
  m = _build(c)
  m.compile(metrics=[MeanSquaredError()])
  out = m.train_step(c.x)
  assert set(out) == {'loss', 'mean_squared_error'}
  assert abs(out['loss'] - c.train_loss) <= 1e-4 * abs(c.train_loss)
  sample = m._staging['sample'].cpu().numpy().reshape(c.B, c.T)
  assert abs(out['mean_squared_error'] - float(((c.x[:, 1:, 0] - sample) ** 2).mean())) < 1e-5
  m.reset_metrics()               # (Keras `evaluate` does; otherwise 'loss' is the running mean over both steps)
  t = m.test_step(c.x)
  assert set(t) == {'loss', 'mean_squared_error'} and abs(t['loss'] - c.test_loss) <= 1e-4 * abs(c.test_loss)
