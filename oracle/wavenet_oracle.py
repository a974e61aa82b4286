"""CPU oracle for the WaveNet residual-stack training pass (TEST INFRASTRUCTURE ONLY).

This file restates, in plain NumPy (fp64 by default, fp32 on request), the arithmetic
of the reference hot path:

  * /root/reference/src/layers.py:10-224   (WaveNetLayer ctor / build / call)
  * /root/reference/src/model.py:9,14-155  (WaveNet ctor: dilation schedule, head, mapping)
  * /root/reference/src/model.py:213-239   (WaveNet.call)
  * /root/reference/src/model.py:309-348   (train_step: shift-by-one, loss scale, gradients)
  * /root/reference/src/model.py:505-551   (loss_fn: categorical / gaussian / logistic)
  * /root/reference/src/utils.py:35-38     (mu-law companding and framing)

with a hand-derived backward pass.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import it; the product (wavenets_b200/) never does.

PARITY PIN: the reference ships no tests, golden vectors or recorded outputs, and its arithmetic
lives in TensorFlow 2 / Keras 3 (un-vendored, unpinned; not installable here).  This restatement is
therefore pinned against tests/golden/*.npz, produced by oracle/make_golden.py from the reference's
OWN source files (src/layers.py, src/model.py, imported unmodified) executed over oracle/tf_shim — a
restatement of only the TF/Keras primitives they call (tests/test_golden_cpu.py).  What stays
restated rather than executed from upstream are those primitives (SURVEY.md section 8c lists them);
they are additionally cross-checked three ways in tests/: this manual backward vs torch.autograd on
an independent torch restatement (oracle/torch_ref.py) vs central finite differences, plus
hand-derived known-answer tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# reference: model.py:9  SQRT2PI = tf.sqrt(2.0*3.14159265359)   (computed in fp32)
SQRT2PI_F32 = float(np.sqrt(np.float32(2.0 * 3.14159265359)))

LEAKY_SLOPE = 0.2  # Keras 3 `activation='leaky_relu'` => negative_slope 0.2


# --------------------------------------------------------------------------- config
@dataclass
class Config:
  """Mirror of WaveNet.__init__ kwargs (model.py:14-34)."""
  kernel_size: int = 2
  channels: int = 32
  blocks: int = 10
  layers_per_block: int = 1
  activation: Optional[str] = None
  conditioning: Optional[str] = None
  mapping_layers: Optional[Sequence[int]] = None
  mapping_activation: Optional[str] = None
  dropout: float = 0.0
  dilation_bound: int = 512
  num_mixtures: Optional[int] = None
  sampling_function: str = 'categorical'
  bits: int = 8
  skip_channels: Optional[int] = None
  dilation_channels: Optional[int] = None
  use_residual: bool = True
  use_skip: bool = True
  final_layers_channels: Sequence[int] = field(default_factory=list)
  l2_reg_factor: float = 0.0
  cond_in: int = 0  # width of the conditioning input vector (one-hot depth)

  def validate(self):
    # model.py:52-70
    if self.conditioning not in ['global', 'local', None]:
      raise ValueError("Conditioning must be 'global', 'local' or None.")
    if self.kernel_size < 2:
      raise ValueError('Kernel size must be at least 2.')
    if math.log(self.dilation_bound, self.kernel_size) % 1 != 0:
      raise ValueError('dilation bound must be power of kernel_size.')
    if self.layers_per_block < 1:
      raise ValueError('Layers per block must be at least 1.')
    if self.blocks < 1:
      raise ValueError('Blocks must be at least 1.')
    if self.num_mixtures is not None and self.num_mixtures < 1:
      raise ValueError('Number of mixtures must be at least 1 or None.')
    if self.dropout < 0 or self.dropout > 1:
      raise ValueError('Dropout must be between 0 and 1.')
    if self.sampling_function not in ['categorical', 'logistic', 'gaussian']:
      raise ValueError('Sampling function must be categorical, logistic or gaussian.')
    if self.sampling_function == 'categorical' and self.num_mixtures is not None:
      raise ValueError('Categorical sampling cannot be used with mixtures.')
    if self.conditioning == 'local':
      raise NotImplementedError('local conditioning is broken in the reference (model.py:137)')

  @property
  def D(self):
    return self.channels if self.dilation_channels is None else self.dilation_channels

  @property
  def out_channels(self):
    return 3 * self.num_mixtures if self.num_mixtures is not None else 2 ** self.bits

  @property
  def cond_channels(self):
    if self.conditioning is None:
      return 0
    ml = list(self.mapping_layers or [])
    return ml[-1] if ml else self.cond_in


def config_from_kwargs(kw: dict, cond_in: int = 0) -> Config:
  """Config from the WaveNet(...) kwargs of model.py:14-34 (what train.py:206-224 passes)."""
  return Config(
    kernel_size=kw.get('kernel_size', 2), channels=kw.get('channels', 32), blocks=kw.get('blocks', 10),
    layers_per_block=kw.get('layers_per_block', 1), activation=kw.get('activation'),
    conditioning=kw.get('conditioning'), mapping_layers=kw.get('mapping_layers'),
    mapping_activation=kw.get('mapping_activation'), dropout=kw.get('dropout', 0.0),
    dilation_bound=kw.get('dilation_bound', 512), num_mixtures=kw.get('num_mixtures'),
    sampling_function=kw.get('sampling_function', 'categorical'), bits=kw.get('bits', 8),
    skip_channels=kw.get('skip_channels'), dilation_channels=kw.get('dilation_channels'),
    use_residual=kw.get('use_residual', True), use_skip=kw.get('use_skip', True),
    final_layers_channels=kw.get('final_layers_channels', []), l2_reg_factor=kw.get('l2_reg_factor', 0.0),
    cond_in=cond_in)


def dilation_schedule(cfg: Config) -> Tuple[List[List[int]], int]:
  """model.py:79-81,93-94,122.  Returns (per-block dilation lists, receptive field)."""
  max_power = int(math.log(cfg.dilation_bound, cfg.kernel_size))
  dil = [cfg.kernel_size ** (i % max_power) for i in range(cfg.layers_per_block * cfg.blocks)]
  per_block = [dil[b * cfg.layers_per_block:(b + 1) * cfg.layers_per_block] for b in range(cfg.blocks)]
  rf = 1 + sum(dil) * (cfg.kernel_size - 1) + 1
  return per_block, rf


def param_specs(cfg: Config) -> List[Tuple[str, Tuple[int, ...]]]:
  """Trainable variables in Keras tracking order (attribute creation order in
  model.py:84-148 / layers.py:64-120), Keras layouts: conv kernel (K,Cin,Cout),
  bias (Cout,), dense kernel (in,out)."""
  K, R, D = cfg.kernel_size, cfg.channels, cfg.D
  S = cfg.skip_channels
  specs = [('causal/kernel', (K, 1, R)), ('causal/bias', (R,))]
  per_block, _ = dilation_schedule(cfg)
  for b, dils in enumerate(per_block):
    cin = R
    for j in range(len(dils) - 1):
      specs += [(f'block{b}/dil{j}/kernel', (K, cin, D)), (f'block{b}/dil{j}/bias', (D,))]
      cin = D
    j = len(dils) - 1
    specs += [(f'block{b}/dil{j}/kernel', (K, cin, 2 * D)), (f'block{b}/dil{j}/bias', (2 * D,))]
    specs += [(f'block{b}/conv1/kernel', (1, D, R)), (f'block{b}/conv1/bias', (R,))]
    if S is not None:
      specs += [(f'block{b}/conv_skip/kernel', (1, D, S)), (f'block{b}/conv_skip/bias', (S,))]
    if cfg.conditioning is not None:
      specs += [(f'block{b}/conv_cond/kernel', (1, cfg.cond_channels, 2 * D)),
                (f'block{b}/conv_cond/bias', (2 * D,))]
  cin = (S if S is not None else R) if cfg.use_skip else R
  for i, ch in enumerate(list(cfg.final_layers_channels) + [cfg.out_channels]):
    specs += [(f'final{i}/kernel', (1, cin, ch)), (f'final{i}/bias', (ch,))]
    cin = ch
  if cfg.conditioning == 'global':
    cin = cfg.cond_in
    for i, ch in enumerate(list(cfg.mapping_layers or [])):
      specs += [(f'mapping{i}/kernel', (cin, ch)), (f'mapping{i}/bias', (ch,))]
      cin = ch
  return specs


def init_params(cfg: Config, seed: int = 1, bias_std: float = 0.02, dtype=np.float64) -> Dict[str, np.ndarray]:
  """Glorot-uniform kernels (Keras default: limit = sqrt(6/(fan_in+fan_out)), fan = receptive
  size * channels).  Biases N(0, bias_std) for parity tests (Keras' zero default would hide
  bias bugs); bias_std=0 reproduces Keras."""
  rng = np.random.default_rng(seed)
  out = {}
  for name, shape in param_specs(cfg):
    if name.endswith('kernel'):
      if len(shape) == 3:
        fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
      else:
        fan_in, fan_out = shape
      lim = math.sqrt(6.0 / (fan_in + fan_out))
      out[name] = rng.uniform(-lim, lim, size=shape).astype(dtype)
    else:
      out[name] = (rng.standard_normal(shape) * bias_std).astype(dtype)
  return out


# --------------------------------------------------------------------------- primitives
def act_fwd(name: Optional[str], x):
  if name is None or name == 'linear':
    return x
  if name == 'relu':
    return np.maximum(x, 0)
  if name == 'leaky_relu':
    return np.where(x >= 0, x, LEAKY_SLOPE * x)
  if name == 'tanh':
    return np.tanh(x)
  if name == 'sigmoid':
    return 1.0 / (1.0 + np.exp(-x))
  raise NotImplementedError(f'activation {name}')


def act_bwd_from_out(name: Optional[str], y, dy):
  """Derivative expressed through the activation OUTPUT y (all supported ones allow it)."""
  if name is None or name == 'linear':
    return dy
  if name == 'relu':
    return dy * (y > 0)
  if name == 'leaky_relu':
    return dy * np.where(y >= 0, 1.0, LEAKY_SLOPE)
  if name == 'tanh':
    return dy * (1.0 - y * y)
  if name == 'sigmoid':
    return dy * y * (1.0 - y)
  raise NotImplementedError(f'activation {name}')


def _shift_right(x, s):
  """y[:, t] = x[:, t-s] with zeros for t-s < 0, PER BATCH ROW (causal left zero padding)."""
  if s == 0:
    return x
  y = np.zeros_like(x)
  if s < x.shape[1]:
    y[:, s:] = x[:, :x.shape[1] - s]
  return y


def _shift_left(x, s):
  """y[:, t] = x[:, t+s] with zeros past the end (adjoint of _shift_right)."""
  if s == 0:
    return x
  y = np.zeros_like(x)
  if s < x.shape[1]:
    y[:, :x.shape[1] - s] = x[:, s:]
  return y


def causal_conv_fwd(x, W, b, dilation):
  """Keras Conv1D(padding='causal'): left-pad d*(K-1) zeros then VALID cross-correlation.
  tap k multiplies x[t - (K-1-k)*d].  x (B,T,Cin), W (K,Cin,Cout), b (Cout,)."""
  K = W.shape[0]
  out = np.zeros(x.shape[:2] + (W.shape[2],), dtype=x.dtype) + b
  for k in range(K):
    out = out + _shift_right(x, (K - 1 - k) * dilation) @ W[k]
  return out


def causal_conv_bwd(x, W, dilation, dout, need_dx=True):
  K = W.shape[0]
  dW = np.zeros_like(W)
  dx = np.zeros_like(x) if need_dx else None
  for k in range(K):
    s = (K - 1 - k) * dilation
    xs = _shift_right(x, s)
    dW[k] = np.einsum('btc,btn->cn', xs, dout)
    if need_dx:
      dx = dx + _shift_left(dout @ W[k].T, s)
  db = dout.sum(axis=(0, 1))
  return dx, dW, db


def discretize(x, bits: int = 8):
  """Keras Discretization(bin_boundaries=linspace(-1,1,2**bits+1)[1:-1]) (model.py:152-153):
  index = number of boundaries <= x, int64.  Comparison based (bit-exact)."""
  bounds = np.asarray(np.linspace(-1, 1, num=2 ** bits + 1).tolist()[1:-1], dtype=np.float32)
  xf = np.asarray(x, dtype=np.float32)
  return np.searchsorted(bounds, xf, side='right').astype(np.int64)


def mu_law(x):
  """utils.py:35 in fp32: sign(x)*log(1+255|x|)/log(256)."""
  x = np.asarray(x, dtype=np.float32)
  return (np.sign(x) * (np.log(np.float32(1.0) + np.float32(255.0) * np.abs(x)) / np.log(np.float32(256.0)))).astype(np.float32)


# --------------------------------------------------------------------------- layer
def layer_forward(p: Dict[str, np.ndarray], prefix: str, cfg_layer: dict, x, cond=None, keep=None, rate=0.0):
  """WaveNetLayer.call (layers.py:178-224).
  cfg_layer: dict(dilations=[...], activation=str|None, residual=bool, has_skip=bool, condition=bool).
  cond: (B,T,Cc) or (B,Cc) (time-constant global conditioning).
  keep/rate: training-mode dropout (layers.py:195-196) with an explicit keep-mask (B,T,R): the conv
  branch sees x*keep/(1-rate), the residual is taken before it (layers.py:192-193).  keep=None: off.
  Returns x_out, skip, cache."""
  dils = cfg_layer['dilations']
  act = cfg_layer['activation'] if len(dils) > 1 else None
  cache = {'x_in': x, 'stack_in': [], 'stack_out': [], 'keep': keep, 'rate': rate}
  h = x
  if keep is not None:
    h = x * keep.astype(x.dtype) / x.dtype.type(1.0 - rate)
  for j, d in enumerate(dils):
    cache['stack_in'].append(h)
    h = causal_conv_fwd(h, p[f'{prefix}/dil{j}/kernel'], p[f'{prefix}/dil{j}/bias'], d)
    if j < len(dils) - 1:
      h = act_fwd(act, h)
    cache['stack_out'].append(h)
  z = h
  if cfg_layer['condition']:
    Wc, bc = p[f'{prefix}/conv_cond/kernel'][0], p[f'{prefix}/conv_cond/bias']
    cz = cond @ Wc + bc
    if cz.ndim == 2:
      cz = cz[:, None, :]
    z = z + cz
  D = z.shape[-1] // 2
  tf_, sg = np.tanh(z[..., :D]), 1.0 / (1.0 + np.exp(-z[..., D:]))
  g = tf_ * sg
  o = g @ p[f'{prefix}/conv1/kernel'][0] + p[f'{prefix}/conv1/bias']
  if cfg_layer['has_skip']:
    skip = g @ p[f'{prefix}/conv_skip/kernel'][0] + p[f'{prefix}/conv_skip/bias']
  else:
    skip = o  # alias taken BEFORE the residual add (layers.py:216-223)
  x_out = o + x if cfg_layer['residual'] else o
  cache.update(tanh=tf_, sig=sg, g=g, cond=cond)
  return x_out, skip, cache


def layer_backward(p, prefix, cfg_layer, cache, dx_out, dskip, need_dx=True):
  """Adjoint of layer_forward.  Returns dx, dcond (or None), grads dict."""
  dils = cfg_layer['dilations']
  act = cfg_layer['activation'] if len(dils) > 1 else None
  grads = {}
  g, tf_, sg = cache['g'], cache['tanh'], cache['sig']
  Wr = p[f'{prefix}/conv1/kernel'][0]
  do = dx_out if dx_out is not None else 0.0
  if not cfg_layer['has_skip'] and dskip is not None:
    do = do + dskip
  if np.isscalar(do):
    do = np.zeros(g.shape[:2] + (Wr.shape[1],), dtype=g.dtype)
  grads[f'{prefix}/conv1/kernel'] = np.einsum('btc,btn->cn', g, do)[None]
  grads[f'{prefix}/conv1/bias'] = do.sum(axis=(0, 1))
  dg = do @ Wr.T
  if cfg_layer['has_skip']:
    Ws = p[f'{prefix}/conv_skip/kernel'][0]
    ds = dskip if dskip is not None else np.zeros(g.shape[:2] + (Ws.shape[1],), dtype=g.dtype)
    grads[f'{prefix}/conv_skip/kernel'] = np.einsum('btc,btn->cn', g, ds)[None]
    grads[f'{prefix}/conv_skip/bias'] = ds.sum(axis=(0, 1))
    dg = dg + ds @ Ws.T
  dz = np.concatenate([dg * sg * (1.0 - tf_ * tf_), dg * tf_ * sg * (1.0 - sg)], axis=-1)
  dcond = None
  if cfg_layer['condition']:
    Wc = p[f'{prefix}/conv_cond/kernel'][0]
    cond = cache['cond']
    if cond.ndim == 2:
      dcz = dz.sum(axis=1)                       # (B,2D)
      grads[f'{prefix}/conv_cond/kernel'] = (cond.T @ dcz)[None]
      grads[f'{prefix}/conv_cond/bias'] = dcz.sum(axis=0)
      dcond = dcz @ Wc.T
    else:
      grads[f'{prefix}/conv_cond/kernel'] = np.einsum('btc,btn->cn', cond, dz)[None]
      grads[f'{prefix}/conv_cond/bias'] = dz.sum(axis=(0, 1))
      dcond = dz @ Wc.T
  dh = dz
  for j in reversed(range(len(dils))):
    if j < len(dils) - 1:
      dh = act_bwd_from_out(act, cache['stack_out'][j], dh)
    W = p[f'{prefix}/dil{j}/kernel']
    dh_in, dW, db = causal_conv_bwd(cache['stack_in'][j], W, dils[j], dh, need_dx=(need_dx or j > 0))
    grads[f'{prefix}/dil{j}/kernel'] = dW
    grads[f'{prefix}/dil{j}/bias'] = db
    dh = dh_in
  dx = dh
  if need_dx and cache.get('keep') is not None:
    dx = dx * cache['keep'].astype(dx.dtype) / dx.dtype.type(1.0 - cache['rate'])
  if need_dx and cfg_layer['residual'] and dx_out is not None:
    dx = dx + dx_out
  return dx, dcond, grads


def _layer_cfgs(cfg: Config):
  per_block, _ = dilation_schedule(cfg)
  return [dict(dilations=d, activation=cfg.activation, residual=cfg.use_residual,
               has_skip=cfg.skip_channels is not None, condition=cfg.conditioning is not None)
          for d in per_block]


# --------------------------------------------------------------------------- model
def mapping_forward(p, cfg: Config, cond_in):
  """model.py:141-148: Dense stack, every layer with mapping_activation, then Identity."""
  h = cond_in
  acts = []
  for i, _ in enumerate(list(cfg.mapping_layers or [])):
    h = act_fwd(cfg.mapping_activation, h @ p[f'mapping{i}/kernel'] + p[f'mapping{i}/bias'])
    acts.append(h)
  return h, acts


def model_forward(p, cfg: Config, x, cond_in=None, return_logits=False, keep_masks=None):
  """WaveNet.call (model.py:213-239).  x (B,T,1); cond_in (B,cond_in).  Returns the model
  output (softmax probabilities or 3M mixture parameters) and a cache for backward.
  keep_masks: per-block dropout keep-masks (training mode, see layer_forward) or None."""
  cache = {'x': x, 'cond_in': cond_in}
  cond = None
  if cfg.conditioning == 'global':
    cond, cache['map_acts'] = mapping_forward(p, cfg, cond_in)
  h = causal_conv_fwd(x, p['causal/kernel'], p['causal/bias'], 1)
  skips, lcaches = [], []
  for b, lc in enumerate(_layer_cfgs(cfg)):
    h, skip, c = layer_forward(p, f'block{b}', lc, h, cond, keep=None if keep_masks is None else keep_masks[b], rate=cfg.dropout)
    skips.append(skip)
    lcaches.append(c)
  if cfg.use_skip:
    h = skips[0]
    for s in skips[1:]:
      h = h + s        # keras.layers.add: left-to-right sum
  cache['layers'] = lcaches
  head_in = [h]
  nfin = len(cfg.final_layers_channels)
  for i in range(nfin):
    h = act_fwd(cfg.activation, h @ p[f'final{i}/kernel'][0] + p[f'final{i}/bias'])
    head_in.append(h)
  logits = h @ p[f'final{nfin}/kernel'][0] + p[f'final{nfin}/bias']
  cache['head_in'] = head_in
  cache['logits'] = logits
  if cfg.num_mixtures is None and not return_logits:
    m = logits.max(axis=-1, keepdims=True)
    e = np.exp(logits - m)
    return e / e.sum(axis=-1, keepdims=True), cache
  return logits, cache


def loss_and_dlogits(cfg: Config, logits, y, scale):
  """loss_fn (model.py:505-551) on the pre-softmax logits / mixture params, per (b,t), and the
  gradient of scale * sum(loss) with respect to `logits`.  y: (B,T,1) float targets
  (x[:,1:,:], model.py:319-320)."""
  dt = logits.dtype
  if cfg.sampling_function == 'categorical':
    # tf.keras.losses.sparse_categorical_crossentropy(target, probs) on the softmax OUTPUT (model.py:114-118,516), Keras 3
    # TF backend [TF-internal, restated; the shim executes the same definition]: c = clip(p, 1e-7, 1 - 1e-7), then sparse
    # softmax cross entropy with log(c) as logits:  l = log(sum_j c_j) - log(c_y).  The clip passes gradient only where
    # 1e-7 <= p <= 1 - 1e-7.  With nothing clipped this is lse - logit_y and softmax - onehot.
    idx = discretize(y[..., 0], cfg.bits)
    eps = 1e-7
    m = logits.max(axis=-1, keepdims=True)
    lse = m[..., 0] + np.log(np.exp(logits - m).sum(axis=-1))
    logp = logits - lse[..., None]
    p = np.exp(logp)
    inside = ((p >= eps) & (p <= 1.0 - eps)).astype(dt)
    c = np.clip(p, eps, 1.0 - eps)
    # sum_j c_j = 1 + sum_j (c_j - p_j): only clipped entries contribute, so an unclipped row gives exactly lse - logit_y
    S = 1.0 + (c - p).sum(axis=-1)
    in_y = np.take_along_axis(inside, idx[..., None], axis=-1)[..., 0]
    logp_y = np.take_along_axis(logp, idx[..., None], axis=-1)[..., 0]
    c_y = np.take_along_axis(c, idx[..., None], axis=-1)[..., 0]
    log_cy = np.where(in_y > 0, logp_y, np.log(c_y))
    loss = np.log(S) - log_cy
    # u_j = dl/dp_j = inside_j (1/S - [j == y]/c_y);  dl/dlogit_k = p_k (u_k - sum_j p_j u_j)
    Pm = (inside * p).sum(axis=-1)
    u = inside / S[..., None]
    np.put_along_axis(u, idx[..., None], np.take_along_axis(u, idx[..., None], -1) - (in_y / c_y)[..., None], axis=-1)
    dot = Pm / S - in_y          # in_y * p_y / c_y == in_y
    d = p * (u - dot[..., None])
    return loss, d * scale
  M = cfg.num_mixtures
  w, mu, ls_raw = logits[..., :M], logits[..., M:2 * M], logits[..., 2 * M:]
  wm = w.max(axis=-1, keepdims=True)
  pi = np.exp(w - wm)
  pi = pi / pi.sum(axis=-1, keepdims=True)
  ls = np.maximum(ls_raw, -7.0)
  pass_ls = (ls_raw >= -7.0).astype(dt)
  yy = y.astype(dt)  # (B,T,1) broadcasts like tf.repeat(target, M, -1)
  if cfg.sampling_function == 'gaussian':
    sig = np.exp(ls)
    xr = (yy - mu) / sig
    xx = np.minimum(xr, 1e8)
    comp = np.exp(-0.5 * xx * xx) / (sig * dt.type(SQRT2PI_F32))
    lik = (pi * comp).sum(axis=-1)
    loss = -np.log(lik)
    r = pi * comp / lik[..., None]
    notclip = (xr <= 1e8).astype(dt)
    dw = pi - r
    dmu = -r * xx / sig * notclip
    dls = -r * (xx * xx * notclip - 1.0) * pass_ls
  else:  # logistic
    h = 0.5 * 1 / (2 ** cfg.bits)
    e = np.exp(-1.0 * ls)
    a, b = (yy - mu + h) * e, (yy - mu - h) * e
    sa, sb = 1.0 / (1.0 + np.exp(-a)), 1.0 / (1.0 + np.exp(-b))
    P = sa - sb
    lik = (pi * P).sum(axis=-1)
    loss = -np.log(lik)
    dsa, dsb = sa * (1 - sa), sb * (1 - sb)
    c = pi / lik[..., None]
    dw = pi - c * P
    dmu = c * e * (dsa - dsb)
    dls = c * (a * dsa - b * dsb) * pass_ls
  return loss, np.concatenate([dw, dmu, dls], axis=-1) * scale


def train_step(p, cfg: Config, x_frames, cond_in=None, n_replicas: int = 1, keep_masks=None):
  """train_step (model.py:309-335) up to the gradients: returns (loss, grads, aux).
  x_frames: (B,T+1,1).  loss = sum_{b,t} l[b,t] / (B * n_replicas)  (compute_average_loss)."""
  cfg.validate()
  dt = x_frames.dtype
  y = x_frames[:, 1:, :]
  inputs = x_frames[:, :-1, :]
  logits, cache = model_forward(p, cfg, inputs, cond_in, return_logits=True, keep_masks=keep_masks)
  B = x_frames.shape[0]
  scale = dt.type(1.0 / (B * n_replicas))
  lpt, dlogits = loss_and_dlogits(cfg, logits, y, scale)
  loss = lpt.sum() * scale
  grads = model_backward(p, cfg, cache, dlogits)
  loss_no_reg = float(loss)
  reg = 0.0
  if cfg.l2_reg_factor > 0:
    # model.py:331-334: reg * sum(w^2) over kernels, scaled by 1/replicas
    for name, _ in param_specs(cfg):
      if name.endswith('kernel'):
        reg += cfg.l2_reg_factor * float((p[name] ** 2).sum())
        grads[name] = grads[name] + 2.0 * cfg.l2_reg_factor * p[name] / n_replicas
    loss = loss + reg / n_replicas
  # the reference reports the two parts separately (metrics 'loss' and 'reg_loss', model.py:340-344)
  return float(loss), grads, {'loss_per_sample': lpt, 'logits': logits, 'dlogits': dlogits, 'cache': cache,
                              'loss_no_reg': loss_no_reg, 'reg_loss': reg / n_replicas}


def clip_by_norm(g, clipnorm):
  """tf.clip_by_norm on one tensor: g * clipnorm / max(||g||_2, clipnorm)."""
  n = float(np.sqrt((np.asarray(g, dtype=np.float64) ** 2).sum()))
  return g * (clipnorm / max(n, clipnorm))


def adam_step(params, grads, state, lr, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=None, n_replicas=1):
  """One `optimizer.apply_gradients` (model.py:336) of tf.keras.optimizers.Adam(lr, clipnorm) as train.py:225-226
  builds it, Keras 3 semantics (restated, TF-internal): per-variable clip_by_norm of each replica's gradient, sum over
  replicas (here: `grads` is one replica's gradient and every replica is assumed identical -> x n_replicas), then
  m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps).
  state: {'t': int, 'm': {...}, 'v': {...}} (created on first use).  Returns (new params, state)."""
  if not state:
    state.update(t=0, m={k: np.zeros_like(v) for k, v in params.items()}, v={k: np.zeros_like(v) for k, v in params.items()})
  state['t'] += 1
  t = state['t']
  alpha = lr * math.sqrt(1.0 - beta_2 ** t) / (1.0 - beta_1 ** t)
  out = {}
  for k, w in params.items():
    g = grads[k]
    if clipnorm:
      g = clip_by_norm(g, clipnorm)
    g = g * n_replicas
    m = state['m'][k] + (g - state['m'][k]) * (1.0 - beta_1)
    v = state['v'][k] + (g * g - state['v'][k]) * (1.0 - beta_2)
    state['m'][k], state['v'][k] = m, v
    out[k] = w - alpha * m / (np.sqrt(v) + epsilon)
  return out, state


def sample_deterministic(cfg: Config, pred):
  """sample_waveform(pred, deterministic=True) (model.py:411-418,452-459,484-499): categorical ->
  argmax bin mapped to [-1,1) as idx/2^(bits-1) - 1, shape (B,T); mixtures -> mean of the heaviest
  component clipped to [-1,1], shape (B,T,1).  pred = model output (probabilities / 3M params)."""
  if cfg.num_mixtures is None:
    idx = np.argmax(pred, axis=-1)
    return idx.astype(pred.dtype) / pred.dtype.type(2.0 ** (cfg.bits - 1)) - pred.dtype.type(1.0)
  M = cfg.num_mixtures
  sel = np.argmax(pred[..., :M], axis=-1)
  mu = np.take_along_axis(pred[..., M:2 * M], sel[..., None], axis=-1)
  return np.clip(mu, -1.0, 1.0)


def model_backward(p, cfg: Config, cache, dlogits):
  grads = {}
  nfin = len(cfg.final_layers_channels)
  head_in = cache['head_in']
  dh = dlogits
  for i in reversed(range(nfin + 1)):
    if i < nfin:
      dh = act_bwd_from_out(cfg.activation, head_in[i + 1], dh)
    W = p[f'final{i}/kernel'][0]
    grads[f'final{i}/kernel'] = np.einsum('btc,btn->cn', head_in[i], dh)[None]
    grads[f'final{i}/bias'] = dh.sum(axis=(0, 1))
    dh = dh @ W.T
  lcfgs = _layer_cfgs(cfg)
  if cfg.use_skip:
    dskip, dx = dh, None
  else:
    dskip, dx = None, dh
  dcond_total = None
  for b in reversed(range(cfg.blocks)):
    dx, dcond, g = layer_backward(p, f'block{b}', lcfgs[b], cache['layers'][b], dx, dskip)
    grads.update(g)
    if dcond is not None:
      dcond_total = dcond if dcond_total is None else dcond_total + dcond
  _, dW, db = causal_conv_bwd(cache['x'], p['causal/kernel'], 1, dx, need_dx=False)
  grads['causal/kernel'], grads['causal/bias'] = dW, db
  if cfg.conditioning == 'global':
    acts = cache['map_acts']
    dh = dcond_total
    ml = list(cfg.mapping_layers or [])
    for i in reversed(range(len(ml))):
      dh = act_bwd_from_out(cfg.mapping_activation, acts[i], dh)
      inp = acts[i - 1] if i > 0 else cache['cond_in']
      grads[f'mapping{i}/kernel'] = inp.T @ dh
      grads[f'mapping{i}/bias'] = dh.sum(axis=0)
      dh = dh @ p[f'mapping{i}/kernel'].T
  return grads


# --------------------------------------------------------------------------- work model
def flops_fwd_per_sample(cfg: Config) -> Dict[str, int]:
  """SURVEY.md 8(d) algorithmic work: F_fwd per audio sample, split by kind."""
  K, R, D = cfg.kernel_size, cfg.channels, cfg.D
  S = cfg.skip_channels
  per_block, _ = dilation_schedule(cfg)
  dil = 0
  one = 0
  for dils in per_block:
    cin = R
    for j in range(len(dils) - 1):
      dil += 2 * K * cin * D
      cin = D
    dil += 2 * K * cin * 2 * D
    one += 2 * D * R + (2 * D * S if S is not None else 0)
  head = 0
  cin = (S if S is not None else R) if cfg.use_skip else R
  for ch in list(cfg.final_layers_channels) + [cfg.out_channels]:
    head += 2 * cin * ch
    cin = ch
  inp = 2 * K * R
  return {'dilated': dil, 'pointwise': one, 'head': head, 'input': inp,
          'total': dil + one + head + inp}


def num_params(cfg: Config) -> int:
  return int(sum(int(np.prod(s)) for _, s in param_specs(cfg)))
