"""Thin Python owner of a `wn_handle` (host-side plumbing only: device memory, streams).

torch is used here exclusively to hold device buffers and CUDA streams; all arithmetic of the
hot path happens inside libwavenet_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib


class _DevArray:
  """Zero-copy torch view of device memory owned by the C library."""

  def __init__(self, ptr: int, n: int, owner):
    self.__cuda_array_interface__ = {
      'shape': (n,), 'typestr': '<f4', 'data': (ptr, False), 'version': 2, 'strides': None}
    self._owner = owner


def _require_cuda():
  if not torch.cuda.is_available():
    raise RuntimeError('wavenets_b200 needs a CUDA device (sm_100a); there is no CPU fallback')


def as_dev(x, device, dtype=torch.float32):
  """numpy / torch (any device) -> contiguous torch CUDA tensor (plumbing)."""
  _require_cuda()
  if isinstance(x, torch.Tensor):
    t = x.detach()
  else:
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
  return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


class Handle:
  def __init__(self, cfg: _lib.WnConfig):
    _require_cuda()
    self.lib = _lib.load()
    self.cfg = cfg
    self.device = torch.device('cuda', cfg.device)
    hp = C.c_void_p()
    _lib.check(self.lib.wn_create(C.byref(cfg), C.byref(hp)))
    self.h = hp
    self.n_params = self.lib.wn_num_params(self.h)
    self.n_scalars = int(self.lib.wn_param_count(self.h))
    self.names: List[str] = []
    self.shapes: List[tuple] = []
    self.offsets: List[int] = []
    name = C.create_string_buffer(128)
    shape = (C.c_int32 * 3)()
    ndim = C.c_int32()
    off = C.c_int64()
    for i in range(self.n_params):
      _lib.check(self.lib.wn_param_info(self.h, i, name, 128, shape, C.byref(ndim), C.byref(off)))
      self.names.append(name.value.decode())
      self.shapes.append(tuple(int(shape[k]) for k in range(ndim.value)))
      self.offsets.append(int(off.value))
    self.index: Dict[str, int] = {n: i for i, n in enumerate(self.names)}
    self.flat_params = torch.as_tensor(_DevArray(self.lib.wn_params_dev(self.h), self.n_scalars, self), device=self.device)
    self.flat_grads = torch.as_tensor(_DevArray(self.lib.wn_grads_dev(self.h), self.n_scalars, self), device=self.device)
    self._loss = torch.zeros(4, dtype=torch.float32, device=self.device)

  def close(self):
    if getattr(self, 'h', None) is not None and self.h.value:
      self.lib.wn_destroy(self.h)
      self.h = C.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  # ---- parameters (Keras layouts)
  def _view(self, flat, i):
    n = int(np.prod(self.shapes[i]))
    return flat[self.offsets[i]:self.offsets[i] + n].view(self.shapes[i])

  def param(self, i_or_name):
    i = self.index[i_or_name] if isinstance(i_or_name, str) else i_or_name
    return self._view(self.flat_params, i)

  def grad(self, i_or_name):
    i = self.index[i_or_name] if isinstance(i_or_name, str) else i_or_name
    return self._view(self.flat_grads, i)

  def set_weights(self, weights: Dict[str, np.ndarray]):
    for name, w in weights.items():
      i = self.index[name]
      a = np.ascontiguousarray(np.asarray(w, dtype=np.float32))
      if tuple(a.shape) != self.shapes[i]:
        raise ValueError(f'{name}: expected shape {self.shapes[i]}, got {a.shape}')
      _lib.check(self.lib.wn_set_param(self.h, i, a.ctypes.data_as(C.c_void_p)))
    self.params_changed()

  def get_weights(self) -> Dict[str, np.ndarray]:
    out = {}
    for i, name in enumerate(self.names):
      a = np.empty(self.shapes[i], dtype=np.float32)
      _lib.check(self.lib.wn_get_param(self.h, i, a.ctypes.data_as(C.c_void_p)))
      out[name] = a
    return out

  def get_grads(self) -> Dict[str, np.ndarray]:
    torch.cuda.synchronize(self.device)
    out = {}
    for i, name in enumerate(self.names):
      a = np.empty(self.shapes[i], dtype=np.float32)
      _lib.check(self.lib.wn_get_grad(self.h, i, a.ctypes.data_as(C.c_void_p)))
      out[name] = a
    return out

  def params_changed(self):
    _lib.check(self.lib.wn_params_changed(self.h, self.stream_ptr()))

  def glorot_init(self, seed: int = 1, bias_std: float = 0.0):
    """Keras defaults: glorot-uniform kernels, zero biases (bias_std>0 for parity tests)."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in zip(self.names, self.shapes):
      if name.endswith('kernel'):
        if len(shape) == 3:
          fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
        else:
          fan_in, fan_out = shape
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
      else:
        w[name] = (rng.standard_normal(shape) * bias_std).astype(np.float32)
    self.set_weights(w)

  # ---- streams
  def stream_ptr(self):
    return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

  @staticmethod
  def ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())
