"""bf16-FAITHFUL oracle mode (TEST INFRASTRUCTURE ONLY; never imported by wavenets_b200/).

Third restatement of the reference path — /root/reference/src/layers.py:178-224 (WaveNetLayer.call),
/root/reference/src/model.py:213-239 (WaveNet.call), :309-335 (train_step up to the gradients), :505-551 (loss_fn) —
on torch CPU float64 with autograd, whose only purpose is to separate bf16 STORAGE error from kernel bugs
(SURVEY.md 8c "bf16-faithful mode"): with `faithful=True` every tensor the CUDA bf16 tier keeps in bf16 is rounded to
bf16 at exactly the point where the kernels round it, and everything the kernels keep in fp32 (TMEM accumulators,
biases, the conditioning path, logits, weight gradients) stays un-rounded (fp64 here: the operands of every product
are bf16 values, so fp64 accumulation is the exact value the fp32 accumulators approximate).

Rounding points of the bf16 tier (wavenets_b200/csrc):
  forward   GEMM weights (packed bf16 copies of the fp32 masters; biases stay fp32); h0 (input conv output); the outputs
            of the pre-stack convs (after the activation); the gate's derivative coefficients P, Q (cached for the backward pass
            INSTEAD of z; they and the gate itself are evaluated on the un-rounded accumulator); g; x_out; the skip sum; the head's hidden activations.
            NOT rounded: the input conv's and the conditioning path's weights and arithmetic (fp32), the logits (fp32).
  backward  d logits; the gradient wrt every head / pre-stack PRE-activation (dgrad epilogue: (dY W^T) * act'(cached bf16
            output)); d skip; d(conv1 output) (= d x_out + d skip when the skip aliases conv1, one rounded sum);
            d z (= d g times the CACHED bf16 coefficients P, Q); d x_out of every block (dgrad + residual gradient, one
            rounded sum); d h0.  NOT rounded: d g (lives in the accumulator), every weight / bias gradient (fp32 sums of
            products / column sums of the rounded tensors above).
With `faithful=False` all roundings are identities and the module is a plain fp64 model (checked against
oracle/torch_ref.py and oracle/wavenet_oracle.py in tests/test_oracle_cpu.py).

Parity pin: same as oracle/wavenet_oracle.py (golden vectors produced by the reference's own source over oracle/tf_shim;
the TF primitives themselves are restated — TensorFlow is not installable here).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import wavenet_oracle as wo
from . import torch_ref


def _bf16(x):
  return x.to(torch.bfloat16).to(x.dtype)


class _RoundFwd(torch.autograd.Function):
  """value stored in bf16; its gradient is NOT stored (lives in an fp32 accumulator)"""

  @staticmethod
  def forward(ctx, x):
    return _bf16(x)

  @staticmethod
  def backward(ctx, g):
    return g


class _RoundBwd(torch.autograd.Function):
  """gradient stored in bf16; the value itself is not rounded here"""

  @staticmethod
  def forward(ctx, x):
    return x.view_as(x)

  @staticmethod
  def backward(ctx, g):
    return _bf16(g)


class _Gate(torch.autograd.Function):
  """g = tanh(z_f) * sigmoid(z_s) on the un-rounded accumulator (layers.py:208-210).  The forward pass caches the gate's two
  derivative coefficients P = dg/dz_f = sig(z_s)(1 - tanh(z_f)^2), Q = dg/dz_s = tanh(z_f) sig(z_s)(1 - sig(z_s)), evaluated
  on the un-rounded accumulator and stored in bf16; the adjoint is d z = [dg P | dg Q], stored in bf16
  (TcEpiGate / TcEpiGateBwd, csrc/tc_epilogues.cuh)."""

  @staticmethod
  def forward(ctx, z):
    D = z.shape[-1] // 2
    th, sg = torch.tanh(z[..., :D]), torch.sigmoid(z[..., D:])
    ctx.save_for_backward(_bf16(sg * (1.0 - th * th)), _bf16(th * sg * (1.0 - sg)))
    return th * sg

  @staticmethod
  def backward(ctx, dg):
    P, Q = ctx.saved_tensors
    return _bf16(torch.cat([dg * P, dg * Q], dim=-1))


class _Act(torch.autograd.Function):
  """y = bf16(act(pre)); d pre = bf16(d y * act'(y)) with act' taken from the cached bf16 output (TcEpiActBwd)."""

  @staticmethod
  def forward(ctx, pre, name, slope_mask=None):
    y = _bf16(torch_ref._act(name, pre))
    ctx.name = name
    ctx.mask = slope_mask
    ctx.save_for_backward(y)
    return y

  @staticmethod
  def backward(ctx, dy):
    (y,) = ctx.saved_tensors
    n = ctx.name
    if n is None or n == 'linear':
      d = dy
    elif n == 'relu':
      d = dy * ((y > 0) if ctx.mask is None else ctx.mask).to(dy.dtype)
    elif n == 'leaky_relu':
      d = dy * torch.where((y >= 0) if ctx.mask is None else ctx.mask, torch.ones_like(y), torch.full_like(y, wo.LEAKY_SLOPE))
    elif n == 'tanh':
      d = dy * (1.0 - y * y)
    elif n == 'sigmoid':
      d = dy * y * (1.0 - y)
    else:
      raise NotImplementedError(n)
    return _bf16(d), None, None


class Rounding:
  """the rounding operators of one mode: identities unless `faithful`"""

  def __init__(self, faithful: bool):
    self.on = faithful

  def fwd(self, x):
    return _RoundFwd.apply(x) if self.on else x

  def bwd(self, x):
    return _RoundBwd.apply(x) if self.on else x

  def both(self, x, tap=None, key=None):
    """`tap[key]` receives the stored tensor: its value is the bf16 value, its .grad (after backward) the bf16 gradient"""
    if not self.on:
      if tap is not None:
        x.retain_grad()
        tap[key] = x
      return x
    t = _RoundFwd.apply(x)
    if tap is not None:
      t.retain_grad()
      tap[key] = t
    return _RoundBwd.apply(t)

  def gate(self, z):
    if self.on:
      return _Gate.apply(z)
    D = z.shape[-1] // 2
    return torch.tanh(z[..., :D]) * torch.sigmoid(z[..., D:])

  def act(self, pre, name, slope_mask=None):
    """slope_mask (bool tensor, optional): which elements take the derivative of the POSITIVE branch of relu / leaky_relu in
    the backward pass, instead of the sign of this oracle's own output.  The derivative of a piecewise-linear activation
    is discontinuous at 0: a pre-activation within one rounding flip of zero lands on different sides in two correct
    implementations, and each such element changes a gradient by a factor 5 (leaky) — with the mask taken from the
    implementation under test the backward arithmetic is compared like for like (the forward values are checked separately)."""
    return _Act.apply(pre, name, slope_mask) if self.on else torch_ref._act(name, pre)


def _conv(x, W, b, d):
  """Keras Conv1D(padding='causal') on channels-last x (B,T,Cin): tap k multiplies x[t - (K-1-k) d], zeros before the
  start of every sequence.  W (K,Cin,Cout)."""
  K, T = W.shape[0], x.shape[1]
  out = None
  for k in range(K):
    s = (K - 1 - k) * d
    xs = x if s == 0 else F.pad(x, (0, 0, s, 0))[:, :T]
    y = xs @ W[k]
    out = y if out is None else out + y
  return out + b


def forward_logits(p, cfg: wo.Config, x, cond_in, R: Rounding, tap=None, slope_masks=None, keep_masks=None):
  """x (B,T,1) -> logits (B,T,C).  tap: optional dict that receives the intermediate tensors the kernels store
  ('h0', ('z', l), ('g', l), ('xout', l), 'skipsum', ('hact', i), 'logits'), each with retain_grad().
  keep_masks: per-block dropout keep-masks (B,T,R) of a training pass (layers.py:195-196: the conv branch sees
  keep * x / (1 - rate), the residual is taken before it, layers.py:192-193); the kernels store that masked input in bf16 (it is
  the operand of the gated conv and of its weight gradient) and keep its gradient in the fp32 accumulator."""
  def keep(key, t):
    if tap is not None:
      t.retain_grad()
      tap[key] = t
    return t

  def mask(key):
    if not slope_masks or key not in slope_masks:
      return None
    return torch.as_tensor(np.asarray(slope_masks[key]), dtype=torch.bool)
  per_block, _ = wo.dilation_schedule(cfg)
  q = lambda name: R.fwd(p[name])     # bf16 copy of a GEMM weight
  cond = None
  if cfg.conditioning == 'global':
    cond = cond_in
    for i, _ in enumerate(list(cfg.mapping_layers or [])):
      cond = torch_ref._act(cfg.mapping_activation, cond @ p[f'mapping{i}/kernel'] + p[f'mapping{i}/bias'])   # fp32 path
  h = R.both(_conv(x, p['causal/kernel'], p['causal/bias'], 1), tap, 'h0')        # element-wise fp32 kernel, stored bf16
  skips = []
  for b, dils in enumerate(per_block):
    res = h
    act = cfg.activation if len(dils) > 1 else None
    a = h
    if keep_masks is not None and cfg.dropout > 0:
      a = R.fwd(h * torch.as_tensor(np.asarray(keep_masks[b]), dtype=h.dtype) / (1.0 - cfg.dropout))
    for j, d in enumerate(dils[:-1]):
      a = keep(('act', b, j), R.act(_conv(a, q(f'block{b}/dil{j}/kernel'), p[f'block{b}/dil{j}/bias'], d), act, mask(('act', b, j))))
    j = len(dils) - 1
    z = _conv(a, q(f'block{b}/dil{j}/kernel'), p[f'block{b}/dil{j}/bias'], dils[-1])
    if cond is not None:
      z = z + (cond @ p[f'block{b}/conv_cond/kernel'][0] + p[f'block{b}/conv_cond/bias'])[:, None, :]
    keep(('z', b), z)
    g = keep(('g', b), R.fwd(R.gate(z)))
    o = R.bwd(g @ q(f'block{b}/conv1/kernel')[0] + p[f'block{b}/conv1/bias'])
    if cfg.skip_channels is not None:
      skip = g @ q(f'block{b}/conv_skip/kernel')[0] + p[f'block{b}/conv_skip/bias']
    else:
      skip = o
    h = R.both(o + res if cfg.use_residual else o, tap, ('xout', b))
    skips.append(skip)
  if cfg.use_skip:
    s = skips[0]
    for t in skips[1:]:
      s = s + t
    h = R.both(s, tap, 'skipsum')
  nfin = len(cfg.final_layers_channels)
  for i in range(nfin + 1):
    lin = h @ q(f'final{i}/kernel')[0] + p[f'final{i}/bias']
    keep(('lin', i), lin)
    h = keep(('hact', i), R.act(lin, cfg.activation, mask(('hact', i)))) if i < nfin else R.bwd(lin)
  return h


def train_step(p_np, cfg: wo.Config, x_frames, cond_in=None, n_replicas=1, faithful=True, threads=None, tap=None, slope_masks=None,
               keep_masks=None):
  """WaveNet.train_step up to the gradients (model.py:309-335).  Returns (loss, grads dict).
  slope_masks: {('hact', i) | ('act', block, j): bool (B,T,C)} — see Rounding.act."""
  if threads:
    torch.set_num_threads(threads)
  R = Rounding(faithful)
  dt = torch.float64
  p = {k: torch.tensor(np.asarray(v), dtype=dt, requires_grad=True) for k, v in p_np.items()}
  x = torch.tensor(np.asarray(x_frames), dtype=dt)
  c = None if cond_in is None else torch.tensor(np.asarray(cond_in), dtype=dt)
  logits = forward_logits(p, cfg, x[:, :-1, :], c, R, tap, slope_masks, keep_masks)
  if faithful:
    logits = logits.to(torch.float32).to(dt)      # the logits live in fp32
  loss = torch_ref.loss_per_sample(cfg, logits, x[:, 1:, :]).sum() / (x.shape[0] * n_replicas)
  total = loss
  if cfg.l2_reg_factor > 0:     # model.py:331-334: reg * sum(kernel^2) / replicas, reported separately ('reg_loss')
    total = loss + cfg.l2_reg_factor * sum((v * v).sum() for k, v in p.items() if k.endswith('kernel')) / n_replicas
  total.backward()
  grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in p.items()}
  return float(loss.detach()), grads


def forward(p_np, cfg: wo.Config, x, cond_in=None, faithful=True):
  """WaveNet.call: probabilities (categorical) or mixture parameters, numpy (B,T,C)."""
  R = Rounding(faithful)
  dt = torch.float64
  with torch.no_grad():
    p = {k: torch.tensor(np.asarray(v), dtype=dt) for k, v in p_np.items()}
    xx = torch.tensor(np.asarray(x), dtype=dt)
    c = None if cond_in is None else torch.tensor(np.asarray(cond_in), dtype=dt)
    logits = forward_logits(p, cfg, xx, c, R)
    if cfg.sampling_function == 'categorical':
      return torch.softmax(logits, dim=-1).numpy()
    return logits.numpy()
