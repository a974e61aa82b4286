"""device input pipeline vs the reference formulas (utils.py:31-70, callbacks.py:126-131)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mu_law_exact(x):
  x = x.astype(np.float64)
  return (np.sign(x) * np.log(1.0 + 255.0 * np.abs(x)) / np.log(256.0)).astype(np.float32)


def _ulp_diff(a, b):
  ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
  return np.abs(ia - ib)


@pytest.mark.parametrize('is_i16', [False, True])
def test_frames_mulaw_filter(is_i16):
  from wavenets_b200 import data
  rng = np.random.default_rng(0)
  T, n = 500, 500 * 7 + 123
  if is_i16:
    speech = rng.integers(-32768, 32767, n).astype(np.int16)
    x = speech.astype(np.float32) / np.float32(32768.0)
  else:
    speech = np.clip(rng.standard_normal(n) * 0.4, -1, 1).astype(np.float32)
    speech[T * 2 + 5] = 1.5            # out of range  -> frame 2 dropped
    speech[T * 4] = np.nan             # shared by frames 3 and 4 (hop T, length T+1) -> both dropped
    x = speech
  frames, cond = data.preprocess_recording(speech, T, True, condition_id=1, condition_depth=2)
  nf = 1 + (n - (T + 1)) // T
  ref = np.stack([_mu_law_exact(x[f * T:f * T + T + 1]) for f in range(nf)])
  keep = np.array([np.isfinite(r).all() and (r >= -1).all() and (r <= 1).all() for r in ref])
  if not is_i16:
    assert list(np.where(~keep)[0]) == [2, 3, 4]
  got = frames.cpu().numpy()[..., 0]
  assert got.shape == (int(keep.sum()), T + 1)
  d = _ulp_diff(got, ref[keep])
  assert d.max() <= 1 and (d == 0).mean() > 0.999          # correctly rounded fp32 of the formula (double rounding aside)
  assert np.array_equal(np.sign(got), np.sign(ref[keep]))
  assert cond.shape == (got.shape[0], 2) and torch.equal(cond[:, 1], torch.ones_like(cond[:, 1])) and float(cond[:, 0].abs().sum()) == 0.0
  # consecutive frames overlap by one sample: the target of the last step of frame f is the input of frame f+1
  if is_i16:
    assert np.array_equal(got[:-1, -1], got[1:, 0])
  # too short for a single frame: empty result, no error
  f0, _ = data.preprocess_recording(speech[:T], T, True)
  assert f0.shape == (0, T + 1, 1)


def test_inverse_mu_law_round_trip_and_one_hot():
  from wavenets_b200 import data
  x = torch.linspace(-1, 1, 100001, device='cuda')
  y, _ = data.preprocess_recording(x, 100000, True)
  back = data.inverse_mu_law(y[0, :, 0])
  assert float((back - x).abs().max()) < 2e-6
  ref = np.sign(y[0, :, 0].cpu().numpy().astype(np.float64)) * (np.power(256.0, np.abs(y[0, :, 0].cpu().numpy().astype(np.float64))) - 1) / 255.0
  assert _ulp_diff(back.cpu().numpy(), ref.astype(np.float32)).max() <= 1
  ids = torch.tensor([0, 3, 108, 5, 200], dtype=torch.int32, device='cuda')
  oh = data.one_hot(ids, 109).cpu().numpy()
  assert oh.shape == (5, 109) and oh.sum() == 4 and oh[1, 3] == 1 and oh[2, 108] == 1 and oh[4].sum() == 0
