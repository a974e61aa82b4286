"""Independent torch-CPU restatement of the reference path (TEST INFRASTRUCTURE ONLY).

Second, independently written restatement of /root/reference/src/layers.py:178-224 and
/root/reference/src/model.py:213-239,309-335,505-551 built on torch CPU ops
(`F.conv1d` with dilation => oneDNN) with autograd for the backward pass.  Two uses:

  * tests/: cross-check of oracle/wavenet_oracle.py's hand-derived backward;
  * bench.py: the `cpu_baseline` / `--impl reference` arm ("CPU restatement of the
    reference path; TensorFlow is not installable here"), multi-threaded.

Parity pin: see the wavenet_oracle.py header (golden vectors from the reference's own source over oracle/tf_shim).  Never imported by wavenets_b200/.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import wavenet_oracle as wo


def _act(name, x):
  if name is None or name == 'linear':
    return x
  if name == 'relu':
    return torch.relu(x)
  if name == 'leaky_relu':
    return F.leaky_relu(x, wo.LEAKY_SLOPE)
  if name == 'tanh':
    return torch.tanh(x)
  if name == 'sigmoid':
    return torch.sigmoid(x)
  raise NotImplementedError(name)


def _causal_conv(x, W, b, d):
  """x (B,C,T) channels-first for torch; W Keras layout (K,Cin,Cout)."""
  K = W.shape[0]
  w = W.permute(2, 1, 0).contiguous()          # (Cout,Cin,K): torch conv1d is cross-correlation too
  return F.conv1d(F.pad(x, (d * (K - 1), 0)), w, b, dilation=d)


def forward_logits(p, cfg: wo.Config, x, cond_in=None):
  """x (B,T,1) -> logits (B,T,C) ; p: dict name -> torch tensor (Keras layouts)."""
  per_block, _ = wo.dilation_schedule(cfg)
  cond = None
  if cfg.conditioning == 'global':
    cond = cond_in
    for i, _ in enumerate(list(cfg.mapping_layers or [])):
      cond = _act(cfg.mapping_activation, cond @ p[f'mapping{i}/kernel'] + p[f'mapping{i}/bias'])
  h = _causal_conv(x.transpose(1, 2), p['causal/kernel'], p['causal/bias'], 1)
  skips = []
  for b, dils in enumerate(per_block):
    res = h
    act = cfg.activation if len(dils) > 1 else None
    for j, d in enumerate(dils):
      h = _causal_conv(h, p[f'block{b}/dil{j}/kernel'], p[f'block{b}/dil{j}/bias'], d)
      if j < len(dils) - 1:
        h = _act(act, h)
    if cond is not None:
      h = h + (cond @ p[f'block{b}/conv_cond/kernel'][0] + p[f'block{b}/conv_cond/bias'])[:, :, None]
    t, s = torch.chunk(h, 2, dim=1)
    g = torch.tanh(t) * torch.sigmoid(s)
    o = _causal_conv(g, p[f'block{b}/conv1/kernel'], p[f'block{b}/conv1/bias'], 1)
    if cfg.skip_channels is not None:
      skip = _causal_conv(g, p[f'block{b}/conv_skip/kernel'], p[f'block{b}/conv_skip/bias'], 1)
    else:
      skip = o
    h = o + res if cfg.use_residual else o
    skips.append(skip)
  if cfg.use_skip:
    h = skips[0]
    for s in skips[1:]:
      h = h + s
  nfin = len(cfg.final_layers_channels)
  for i in range(nfin + 1):
    h = _causal_conv(h, p[f'final{i}/kernel'], p[f'final{i}/bias'], 1)
    if i < nfin:
      h = _act(cfg.activation, h)
  return h.transpose(1, 2)


def loss_per_sample(cfg: wo.Config, logits, y):
  """model.py:505-551 on logits; y (B,T,1)."""
  if cfg.sampling_function == 'categorical':
    # sparse_categorical_crossentropy on the softmax output, Keras 3 [TF-internal, restated]: clip(p, 1e-7, 1 - 1e-7), log,
    # then sparse softmax cross entropy on those "logits" (= plain cross entropy on the logits while nothing is clipped)
    idx = torch.from_numpy(wo.discretize(y[..., 0].detach().numpy(), cfg.bits))
    logp = torch.log_softmax(logits, dim=-1)
    pr = torch.exp(logp)
    inside = (pr >= 1e-7) & (pr <= 1.0 - 1e-7)
    c = torch.clamp(pr, 1e-7, 1.0 - 1e-7)
    logc = torch.where(inside, logp, torch.log(c))           # log(p) evaluated as logit - lse where the clip is inactive
    lognorm = torch.log1p((c - pr).sum(-1))                  # log(sum_j c_j), sum_j p_j == 1
    return lognorm - torch.gather(logc, -1, idx.long().unsqueeze(-1)).squeeze(-1)
  w, mu, ls = torch.chunk(logits, 3, dim=-1)
  pi = torch.softmax(w, dim=-1)
  ls = torch.clamp_min(ls, -7.0)
  if cfg.sampling_function == 'gaussian':
    sc = torch.exp(ls)
    xx = torch.clamp_max((y - mu) / sc, 1e8)
    lik = (pi * (torch.exp(-0.5 * xx * xx) / (sc * wo.SQRT2PI_F32))).sum(-1)
  else:
    h = 0.5 * 1 / (2 ** cfg.bits)
    lik = (pi * (torch.sigmoid((y - mu + h) * torch.exp(-1.0 * ls))
                 - torch.sigmoid((y - mu - h) * torch.exp(-1.0 * ls)))).sum(-1)
  return -1.0 * torch.log(lik)


def train_step(p_np, cfg: wo.Config, x_frames, cond_in=None, n_replicas=1, dtype=torch.float64):
  """Returns (loss, grads dict of numpy arrays)."""
  p = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in p_np.items()}
  x = torch.tensor(np.asarray(x_frames), dtype=dtype)
  c = None if cond_in is None else torch.tensor(np.asarray(cond_in), dtype=dtype)
  logits = forward_logits(p, cfg, x[:, :-1, :], c)
  loss = loss_per_sample(cfg, logits, x[:, 1:, :]).sum() / (x.shape[0] * n_replicas)
  loss.backward()
  grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in p.items()}
  return float(loss.detach()), grads


class CpuStepper:
  """fwd+bwd stepper for the CPU baseline: fp32, all host threads."""

  def __init__(self, cfg: wo.Config, p_np, threads=None):
    self.cfg = cfg
    if threads:
      torch.set_num_threads(threads)
    self.p = {k: torch.tensor(np.asarray(v), dtype=torch.float32, requires_grad=True) for k, v in p_np.items()}

  def step(self, x_frames, cond_in=None):
    for v in self.p.values():
      v.grad = None
    x = torch.as_tensor(x_frames, dtype=torch.float32)
    c = None if cond_in is None else torch.as_tensor(cond_in, dtype=torch.float32)
    logits = forward_logits(self.p, self.cfg, x[:, :-1, :], c)
    loss = loss_per_sample(self.cfg, logits, x[:, 1:, :]).sum() / x.shape[0]
    loss.backward()
    return float(loss.detach())
