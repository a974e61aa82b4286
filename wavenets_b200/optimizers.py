"""Adam with per-variable clipnorm — stands where `tf.keras.optimizers.Adam(learning_rate=lr, clipnorm=1.0)`
does in the reference (train.py:225-226, applied by model.py:336).  Keras 3 defaults: beta_1 0.9,
beta_2 0.999, epsilon 1e-7.  The update runs in libwavenet_b200.so on the flat fp32 master weights;
`learning_rate` is a plain attribute so `ReduceLROnPlateau`-style schedules can assign to it
(train.py:167-171)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class Adam:
  def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7,
               clipnorm=None, **kwargs):
    if kwargs.get('clipvalue') is not None or kwargs.get('global_clipnorm') is not None:
      raise NotImplementedError('only clipnorm is built (the reference uses clipnorm=1.0)')
    self.learning_rate = float(learning_rate)
    self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
    self.clipnorm = None if clipnorm is None else float(clipnorm)
    self._handle = None
    self.iterations = 0

  # model.py:211 `self.optimizer.build(self.trainable_variables)`
  def build(self, model_or_variables=None):
    model = model_or_variables
    if model is None or not hasattr(model, 'handle'):
      return
    h = model.handle
    if self._handle is not h:
      _lib.check(h.lib.wn_adam_init(h.h, self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self.clipnorm or 0.0))
      self._handle = h
      self.iterations = 0

  def clip(self, model):
    """Per-variable tf.clip_by_norm in place; Keras does this on every replica before the gradient all-reduce."""
    self.build(model)
    h = model.handle
    _lib.check(h.lib.wn_clip_grads(h.h, h.stream_ptr()))

  def apply_gradients(self, model):
    """One update from `model.handle.flat_grads` (already clipped and, multi-GPU, all-reduced by train_step)."""
    self.build(model)
    h = model.handle
    _lib.check(h.lib.wn_adam_step(h.h, float(self.learning_rate), h.stream_ptr()))
    self.iterations += 1

  def grad_norms(self, model):
    """L2 norm of every variable's gradient as seen by the last clip() (before clipping)."""
    h = model.handle
    p = C.c_void_p()
    _lib.check(h.lib.wn_adam_state(h.h, None, None, C.byref(p), None))
    buf = torch.empty(h.n_params, dtype=torch.float32, device=h.device)
    torch.cuda.current_stream(h.device).synchronize()
    src = torch.as_tensor(_DevView(p.value, h.n_params), device=h.device)
    buf.copy_(src)
    return dict(zip(h.names, buf.cpu().numpy().tolist()))

  def get_state(self, model):
    """First / second moments in Keras layouts (for tests and checkpoints)."""
    h = model.handle
    m, v = C.c_void_p(), C.c_void_p()
    _lib.check(h.lib.wn_adam_state(h.h, C.byref(m), C.byref(v), None, None))
    out = {}
    for name, ptr in (('m', m.value), ('v', v.value)):
      flat = torch.as_tensor(_DevView(ptr, h.n_scalars), device=h.device).cpu().numpy()
      out[name] = {n: flat[o:o + int(np.prod(s))].reshape(s).copy() for n, s, o in zip(h.names, h.shapes, h.offsets)}
    return out


class _DevView:
  def __init__(self, ptr: int, n: int):
    self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f4', 'data': (ptr, False), 'version': 2, 'strides': None}
