"""Device-side input pipeline — stands where `src/utils.py:preprocess_dataset` (utils.py:22-85) does: int16 -> float,
optional mu-law companding (utils.py:35), framing into recording_length+1 samples with hop recording_length
(utils.py:36-38), the finite / range filter (utils.py:58-70) and the one-hot condition (utils.py:47); plus
`inverse_mu_law` (callbacks.py:126-131).  The arithmetic runs in libwavenet_b200.so; torch only holds the buffers."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream(dev):
  return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def preprocess_recording(speech, recording_length: int, apply_mulaw: bool, condition_id=None, condition_depth: int = 2, device=0):
  """One recording -> (frames (n, recording_length+1, 1) fp32, cond (n, depth) or None), invalid frames dropped.
  `speech`: 1-D int16 (scaled by 2^-15 like utils.py:52-55) or float array / tensor."""
  if not torch.cuda.is_available():
    raise RuntimeError('wavenets_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
  lib = _lib.load()
  dev = torch.device('cuda', device)
  t = speech if isinstance(speech, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(speech))
  is_i16 = t.dtype == torch.int16
  t = t.to(device=dev, dtype=torch.int16 if is_i16 else torch.float32).contiguous().flatten()
  T = int(recording_length)
  n = int(lib.wn_num_frames(t.numel(), T))
  frames = torch.empty((n, T + 1), dtype=torch.float32, device=dev)
  valid = torch.empty((n,), dtype=torch.int32, device=dev)
  _lib.check(lib.wn_preprocess_frames(C.c_void_p(t.data_ptr()), 1 if is_i16 else 0, t.numel(), T, 1 if apply_mulaw else 0,
                                      C.c_void_p(frames.data_ptr()), C.c_void_p(valid.data_ptr()), _stream(dev)))
  keep = valid.bool()
  frames = frames[keep].unsqueeze(-1)
  if condition_id is None:
    return frames, None
  ids = torch.full((frames.shape[0],), int(condition_id), dtype=torch.int32, device=dev)
  return frames, one_hot(ids, condition_depth)


def one_hot(ids, depth: int):
  lib = _lib.load()
  ids = ids.to(dtype=torch.int32).contiguous()
  out = torch.empty((ids.numel(), int(depth)), dtype=torch.float32, device=ids.device)
  _lib.check(lib.wn_one_hot(C.c_void_p(ids.data_ptr()), ids.numel(), int(depth), C.c_void_p(out.data_ptr()), _stream(ids.device)))
  return out


def inverse_mu_law(y):
  """callbacks.py:126-131: sign(y) * (256^|y| - 1) / 255."""
  lib = _lib.load()
  y = y.to(dtype=torch.float32).contiguous()
  x = torch.empty_like(y)
  _lib.check(lib.wn_inverse_mu_law(C.c_void_p(y.data_ptr()), C.c_void_p(x.data_ptr()), y.numel(), _stream(y.device)))
  return x
