"""Synthetic inputs of the benchmark: mu-law audio frames and speaker ids.

Host-side data generation only (numpy); the formulas follow the reference input pipeline:
mu-law companding utils.py:35, frames of recording_length+1 samples utils.py:36-38,
one-hot conditioning vector utils.py:47 (109 VCTK speakers instead of 2 genders).
"""
from __future__ import annotations

import numpy as np


def mu_law(x: np.ndarray) -> np.ndarray:
  """utils.py:35 — sign(x) * log(1 + 255|x|) / log(256), fp32."""
  x = np.asarray(x, dtype=np.float32)
  return (np.sign(x) * (np.log(np.float32(1.0) + np.float32(255.0) * np.abs(x)) / np.log(np.float32(256.0)))).astype(np.float32)


def frames(batch: int, recording_length: int, seed: int = 0, fs: int = 16000, apply_mulaw: bool = True) -> np.ndarray:
  """(B, T+1, 1) fp32 in [-1,1]: three sinusoids (80..4000 Hz) * 0.3 + N(0, 0.05), clipped."""
  rng = np.random.default_rng(seed)
  t = np.arange(recording_length + 1, dtype=np.float64) / fs
  x = np.zeros((batch, recording_length + 1), dtype=np.float64)
  for b in range(batch):
    for _ in range(3):
      f = rng.uniform(80.0, 4000.0)
      ph = rng.uniform(0.0, 2 * np.pi)
      x[b] += 0.3 * np.sin(2 * np.pi * f * t + ph)
    x[b] += rng.normal(0.0, 0.05, size=t.shape)
  x = np.clip(x, -1.0, 1.0).astype(np.float32)
  if apply_mulaw:
    x = mu_law(x)
  return x[:, :, None]


def speakers_onehot(batch: int, n_speakers: int = 109, seed: int = 0) -> np.ndarray:
  rng = np.random.default_rng(seed + 12345)
  ids = rng.integers(0, n_speakers, batch)
  out = np.zeros((batch, n_speakers), dtype=np.float32)
  out[np.arange(batch), ids] = 1.0
  return out
