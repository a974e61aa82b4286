"""Minimal stand-in for the `tensorflow` package (TEST INFRASTRUCTURE ONLY — lives under oracle/).

Purpose: run the reference's OWN source files (/root/reference/src/layers.py, src/model.py),
unmodified, in a container where TensorFlow cannot be installed, so that the golden vectors under
tests/golden/ are produced by the reference's code (layer wiring, dilation schedule, split order,
slicing, loss formulas, regularisation, loss scaling) rather than by a re-reading of it.  Only
the TF/Keras PRIMITIVES the reference calls are restated here, on torch CPU tensors (float64 by
default), each from its published definition:

  Conv1D            Keras 3: padding='causal' left-pads dilation*(kernel-1) zeros, then VALID
                    cross-correlation; kernel (K, Cin, Cout), bias (Cout); glorot-uniform / zeros
  Dense             y = act(x @ kernel + bias), kernel (in, out)
  Dropout           inverted dropout, identity when training is False
  Discretization    tf Bucketize: index = number of boundaries <= x, int64
  activations       'leaky_relu' => negative_slope 0.2 (Keras 3), relu, tanh, sigmoid, softmax, linear
  regularizers.L2   l2 * sum(w^2)
  sparse_categorical_crossentropy (from_logits=False), Keras 3 TF backend: clip(p, 1e-7, 1-1e-7),
                    log, then sparse softmax cross entropy on those "logits"
  nn.compute_average_loss  sum(per_example) / (per_example.shape[0] * num_replicas)
  nn.scale_regularization_loss  x / num_replicas
  GradientTape      torch.autograd

`set_num_replicas(n)` emulates `tf.distribute` replica count for the two loss-scaling calls.
Sampling ops (tf.random.*) are implemented only so that train_step runs to the end; their
outputs are never used as golden values (TF's RNG streams are not reproducible here).
"""
from __future__ import annotations

import math as _math
import types

import numpy as np
import torch

__version__ = '0.0-shim'

_DTYPE = torch.float64
_NUM_REPLICAS = 1


def set_default_dtype(dt):
  global _DTYPE
  _DTYPE = dt


def set_num_replicas(n):
  global _NUM_REPLICAS
  _NUM_REPLICAS = int(n)


float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
Tensor = torch.Tensor


def _t(x, dtype=None):
  if isinstance(x, Variable):
    return x.value
  if isinstance(x, torch.Tensor):
    return x
  a = np.asarray(x)
  if a.dtype.kind == 'f':
    return torch.as_tensor(a, dtype=dtype or _DTYPE)
  return torch.as_tensor(a)


class Variable:
  def __init__(self, value, name=None, trainable=True):
    self.value = torch.as_tensor(value).clone().detach().requires_grad_(trainable)
    self.name = name
    self.trainable = trainable

  @property
  def shape(self):
    return tuple(self.value.shape)

  def assign(self, v):
    with torch.no_grad():
      self.value.copy_(torch.as_tensor(np.asarray(v), dtype=self.value.dtype).reshape(self.value.shape))

  def numpy(self):
    return self.value.detach().numpy()


class TensorSpec:
  def __init__(self, shape=None, dtype=None, name=None):
    self.shape, self.dtype, self.name = shape, dtype, name


def function(fn=None, **kwargs):
  """@tf.function and @tf.function(input_signature=...): eager execution, no tracing."""
  if fn is None:
    return lambda f: f
  return fn


# ------------------------------------------------------------------ math
def sqrt(x):
  if isinstance(x, (int, float)):   # tf.sqrt(python float) -> float32 constant (model.py:9)
    return torch.tensor(float(np.sqrt(np.float32(x))), dtype=torch.float32)
  return torch.sqrt(_t(x))


def split(x, n, axis=-1):
  x = _t(x)
  assert x.shape[axis] % n == 0
  return list(torch.chunk(x, n, dim=axis))


def exp(x): return torch.exp(_t(x))
def square(x): return _t(x) ** 2
def maximum(a, b): return torch.maximum(_t(a), torch.as_tensor(b, dtype=_t(a).dtype))
def minimum(a, b): return torch.minimum(_t(a), torch.as_tensor(b, dtype=_t(a).dtype))
def expand_dims(x, axis): return torch.unsqueeze(_t(x), axis)
def squeeze(x, axis=None): return torch.squeeze(_t(x)) if axis is None else torch.squeeze(_t(x), axis)
def concat(xs, axis): return torch.cat([_t(x) for x in xs], dim=axis)
def repeat(x, repeats, axis): return torch.repeat_interleave(_t(x), int(repeats), dim=axis)
def zeros(shape, dtype=None): return torch.zeros(tuple(int(s) for s in shape), dtype=dtype or _DTYPE)
def shape(x): return torch.tensor(tuple(_t(x).shape))
def cast(x, dtype): return _t(x).to(dtype if dtype not in (float32,) else _DTYPE)
def argmax(x, axis=-1): return torch.argmax(_t(x), dim=axis)
def one_hot(idx, depth): return torch.nn.functional.one_hot(_t(idx).long(), int(depth)).to(_DTYPE)
def clip_by_value(x, lo, hi): return torch.clamp(_t(x), lo, hi)


def reduce_sum(x, axis=None):
  if isinstance(x, (list, tuple)):
    if len(x) == 0:
      return torch.zeros((), dtype=_DTYPE)
    x = torch.stack([_t(v) for v in x])
  x = _t(x)
  return x.sum() if axis is None else x.sum(dim=axis)


def map_fn(fn, elems):
  return torch.stack([fn(e) for e in _t(elems)])


math_ns = types.SimpleNamespace(tanh=lambda x: torch.tanh(_t(x)), sigmoid=lambda x: torch.sigmoid(_t(x)),
                                log=lambda x: torch.log(_t(x)), exp=exp, square=square)
# `tf.math`
globals()['math'] = math_ns


# ------------------------------------------------------------------ nn
def _conv1d_valid(x, kernel, dilation=1):
  """x (B,T,Cin), kernel (K,Cin,Cout): VALID cross-correlation along T."""
  w = kernel.permute(2, 1, 0)              # (Cout, Cin, K)
  y = torch.nn.functional.conv1d(x.transpose(1, 2), w, dilation=dilation)
  return y.transpose(1, 2)


def _nn_conv1d(x, kernel, stride=1, padding='VALID'):
  assert stride == 1 and padding == 'VALID'
  return _conv1d_valid(_t(x), _t(kernel))


class _ImmutableTensor(torch.Tensor):
  """tf.Tensor is immutable: `a += b` rebinds `a` (model.py:333-334 relies on it: `loss_final = loss;
  loss_final += reg_loss` must leave `loss` untouched).  torch's `+=` would write in place."""

  def __iadd__(self, o): return self + o
  def __isub__(self, o): return self - o
  def __imul__(self, o): return self * o
  def __itruediv__(self, o): return self / o


def _compute_average_loss(per_example_loss, global_batch_size=None):
  p = _t(per_example_loss)
  if global_batch_size is None:
    global_batch_size = p.shape[0] * _NUM_REPLICAS
  return (p.sum() / global_batch_size).as_subclass(_ImmutableTensor)


nn = types.SimpleNamespace(
  softmax=lambda x, axis=-1: torch.softmax(_t(x), dim=axis),
  sigmoid=lambda x: torch.sigmoid(_t(x)),
  conv1d=_nn_conv1d,
  bias_add=lambda x, b: _t(x) + _t(b),
  compute_average_loss=_compute_average_loss,
  scale_regularization_loss=lambda x: _t(x) / _NUM_REPLICAS,
)


# ------------------------------------------------------------------ random (never used for golden values)
def _gen(seed):
  g = torch.Generator()
  g.manual_seed(int(seed[0]) * 1000003 + int(seed[1]))
  return g


def _shape_tuple(s):
  return tuple(int(v) for v in (s.tolist() if isinstance(s, torch.Tensor) else s))


random = types.SimpleNamespace(
  stateless_categorical=lambda logits, n, seed, dtype=None: torch.multinomial(
    torch.softmax(_t(logits), dim=-1), n, replacement=True, generator=_gen(seed)),
  stateless_normal=lambda shape, seed, dtype=None: torch.randn(_shape_tuple(shape), generator=_gen(seed), dtype=_DTYPE),
  stateless_uniform=lambda shape, seed, dtype=None: torch.rand(_shape_tuple(shape), generator=_gen(seed), dtype=_DTYPE),
)


# ------------------------------------------------------------------ autodiff
class GradientTape:
  def __enter__(self):
    return self

  def __exit__(self, *a):
    return False

  def gradient(self, target, sources):
    src = [s.value if isinstance(s, Variable) else s for s in sources]
    g = torch.autograd.grad(target, src, allow_unused=True)
    return [torch.zeros_like(s) if gi is None else gi for gi, s in zip(g, src)]


# ------------------------------------------------------------------ keras
_ACT = {
  None: lambda x: x, 'linear': lambda x: x,
  'relu': torch.relu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid,
  'leaky_relu': lambda x: torch.nn.functional.leaky_relu(x, 0.2),   # Keras 3 default negative_slope
  'softmax': lambda x: torch.softmax(x, dim=-1),
}


class _L2:
  def __init__(self, l2=0.01):
    self.l2 = 0.0 if l2 is None else float(l2)

  def __call__(self, w):
    return self.l2 * (_t(w) ** 2).sum()


class _Tracked:
  """Attribute-order tracking of sub-layers / variables / lists of layers (Keras' auto-tracking)."""

  def __setattr__(self, k, v):
    order = self.__dict__.setdefault('_tracked_order', [])
    if isinstance(v, (Layer, Variable, list)) and k not in order:
      order.append(k)
    object.__setattr__(self, k, v)

  def _children(self):
    for k in self.__dict__.get('_tracked_order', []):
      v = self.__dict__.get(k)
      if isinstance(v, list):
        for e in v:
          if isinstance(e, (Layer, Variable)):
            yield e
      elif isinstance(v, (Layer, Variable)):
        yield v

  @property
  def trainable_variables(self):
    out = []
    for c in self._children():
      if isinstance(c, Variable):
        if c.trainable:
          out.append(c)
      else:
        out.extend(c.trainable_variables)
    return out

  @property
  def losses(self):
    out = []
    for c in self._children():
      if isinstance(c, Layer):
        out.extend(c._own_losses())
        out.extend(c.losses)
    return out


class Layer(_Tracked):
  def __init__(self, **kwargs):
    self.built = False
    self.name = kwargs.get('name', type(self).__name__.lower())

  def build(self, input_shape):
    self.built = True

  def _own_losses(self):
    return []

  def compute_output_shape(self, input_shape):
    return input_shape

  def __call__(self, inputs, *args, **kwargs):
    if not self.built:
      self.build(_shape_of(inputs))
      self.built = True
    # Keras propagates `training` through the call context to layers called without it (layers.py:196)
    pushed = 'training' in kwargs and kwargs['training'] is not None
    if pushed:
      _TRAINING.append(bool(kwargs['training']))
    try:
      return self.call(inputs, *args, **kwargs)
    finally:
      if pushed:
        _TRAINING.pop()


_TRAINING = [False]


def _shape_of(x):
  if isinstance(x, (list, tuple)):
    return [_shape_of(v) for v in x]
  return tuple(_t(x).shape)


_INIT_GEN = torch.Generator().manual_seed(1234)


def _glorot(shape, fan_in, fan_out):
  lim = _math.sqrt(6.0 / (fan_in + fan_out))
  return (torch.rand(shape, generator=_INIT_GEN, dtype=torch.float64) * 2 - 1).to(_DTYPE) * lim


class Conv1D(Layer):
  def __init__(self, filters, kernel_size, strides=1, padding='valid', dilation_rate=1, activation=None,
               kernel_regularizer=None, **kwargs):
    super().__init__(**kwargs)
    assert strides == 1
    self.filters, self.kernel_size, self.padding = int(filters), int(kernel_size), padding
    self.dilation_rate, self.activation, self.kernel_regularizer = int(dilation_rate), activation, kernel_regularizer
    if activation not in _ACT:
      raise ValueError(f'unknown activation {activation!r}')
    if padding not in ('valid', 'same', 'causal'):
      raise ValueError(f'unknown padding {padding!r}')

  def build(self, input_shape):
    cin = int(input_shape[-1])
    k = self.kernel_size
    self.kernel = Variable(_glorot((k, cin, self.filters), k * cin, k * self.filters), 'kernel')
    self.bias = Variable(torch.zeros(self.filters, dtype=_DTYPE), 'bias')
    self.built = True

  def _own_losses(self):
    return [self.kernel_regularizer(self.kernel)] if (self.kernel_regularizer is not None and self.built) else []

  def compute_output_shape(self, input_shape):
    t = input_shape[1]
    if self.padding == 'valid' and t is not None:
      t = t - self.dilation_rate * (self.kernel_size - 1)
    return (input_shape[0], t, self.filters)

  def call(self, x):
    x = _t(x)
    span = self.dilation_rate * (self.kernel_size - 1)
    if self.padding == 'causal':
      x = torch.nn.functional.pad(x, (0, 0, span, 0))
    elif self.padding == 'same':
      lo = span // 2
      x = torch.nn.functional.pad(x, (0, 0, lo, span - lo))
    y = _conv1d_valid(x, self.kernel.value, self.dilation_rate) + self.bias.value
    return _ACT[self.activation](y)


class Dense(Layer):
  def __init__(self, units, activation=None, kernel_regularizer=None, **kwargs):
    super().__init__(**kwargs)
    self.units, self.activation, self.kernel_regularizer = int(units), activation, kernel_regularizer

  def build(self, input_shape):
    cin = int(input_shape[-1])
    self.kernel = Variable(_glorot((cin, self.units), cin, self.units), 'kernel')
    self.bias = Variable(torch.zeros(self.units, dtype=_DTYPE), 'bias')
    self.built = True

  def _own_losses(self):
    return [self.kernel_regularizer(self.kernel)] if (self.kernel_regularizer is not None and self.built) else []

  def compute_output_shape(self, input_shape):
    return tuple(input_shape[:-1]) + (self.units,)

  def call(self, x):
    return _ACT[self.activation](_t(x) @ self.kernel.value + self.bias.value)


class Dropout(Layer):
  def __init__(self, rate, **kwargs):
    super().__init__(**kwargs)
    self.rate = float(rate)
    self.mask = None          # test hook: inject a fixed keep-mask (TF's RNG stream is not reproducible)

  def call(self, x, training=None):
    x = _t(x)
    if training is None:
      training = _TRAINING[-1]
    if not training or self.rate == 0:
      return x
    keep = self.mask if self.mask is not None else (torch.rand(x.shape) >= self.rate)
    return x * keep.to(x.dtype) / (1.0 - self.rate)


class Identity(Layer):
  def call(self, x):
    return _t(x)


class Lambda(Layer):
  def __init__(self, fn, arguments=None, **kwargs):
    super().__init__(**kwargs)
    self.fn, self.arguments = fn, arguments or {}

  def call(self, x):
    return self.fn(x, **self.arguments)


class Discretization(Layer):
  def __init__(self, bin_boundaries, **kwargs):
    super().__init__(**kwargs)
    self.bin_boundaries = torch.tensor(list(bin_boundaries), dtype=torch.float64)

  def call(self, x):
    x = _t(x).to(torch.float64)
    return torch.searchsorted(self.bin_boundaries, x.contiguous(), right=True)   # count of boundaries <= x


def _layers_add(xs):
  out = _t(xs[0])
  for v in xs[1:]:
    out = out + _t(v)
  return out


class Sequential(Layer):
  def __init__(self, layers=None, **kwargs):
    super().__init__(**kwargs)
    self.layers = list(layers or [])

  def add(self, layer):
    self.layers.append(layer)

  def build(self, input_shape):
    s = input_shape
    for l in self.layers:
      if not l.built:
        l.build(s)
        l.built = True
      s = l.compute_output_shape(s)
    self.built = True

  def call(self, x):
    for l in self.layers:
      x = l(x)
    return x


class Model(Layer):
  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    self.optimizer = None

  def compile(self, optimizer=None, metrics=None, **kwargs):
    object.__setattr__(self, 'optimizer', optimizer)

  def __call__(self, inputs, training=False):
    if not self.built:
      self.build(_shape_of(inputs))
    _TRAINING.append(bool(training))
    try:
      return self.call(inputs, training=training)
    finally:
      _TRAINING.pop()


class _Mean:
  def __init__(self, name='mean'):
    self.name, self.total, self.count = name, 0.0, 0

  def update_state(self, v, *a):
    self.total += float(_t(v).detach().mean())
    self.count += 1

  def result(self):
    return self.total / max(1, self.count)


class RecordingOptimizer:
  """Stands where `tf.keras.optimizers.Adam` would: records what train_step hands to
  `apply_gradients` (model.py:336) and, optionally, applies a caller-supplied update rule."""

  def __init__(self, update=None):
    self.gradients, self.variables, self.update, self.built_with = None, None, update, None

  def build(self, variables):
    self.built_with = list(variables)

  def apply_gradients(self, grads_and_vars):
    gv = list(grads_and_vars)
    self.gradients = [g.detach().clone() for g, _ in gv]
    self.variables = [v for _, v in gv]
    if self.update is not None:
      self.update(self.gradients, self.variables)


def _sparse_categorical_crossentropy(target, output, from_logits=False, axis=-1):
  output = _t(output)
  target = _t(target)
  if target.dim() == output.dim() and target.shape[-1] == 1:
    target = target.squeeze(-1)
  if not from_logits:
    eps = 1e-7
    output = torch.log(torch.clamp(output, eps, 1.0 - eps))
  logp = torch.log_softmax(output, dim=-1)
  return -torch.gather(logp, -1, target.long().unsqueeze(-1)).squeeze(-1)


keras = types.SimpleNamespace(
  Model=Model, Sequential=Sequential,
  layers=types.SimpleNamespace(Layer=Layer, Conv1D=Conv1D, Dense=Dense, Dropout=Dropout, Identity=Identity, Lambda=Lambda,
                               Discretization=Discretization, add=_layers_add),
  regularizers=types.SimpleNamespace(L2=_L2),
  metrics=types.SimpleNamespace(Mean=_Mean),
  losses=types.SimpleNamespace(sparse_categorical_crossentropy=_sparse_categorical_crossentropy),
  activations=types.SimpleNamespace(),
)
