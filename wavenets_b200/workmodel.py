"""Algorithmic work of the training pass per audio sample: the figures the
roofline numbers of bench.py are computed from.  Pure arithmetic on the WaveNet(...) kwargs of
model.py:14-34 — no compute, no device.

  F_fwd = 2 K R + sum_blocks[ sum_{i<last} 2 K Cin_i D + 2 K Cin_last 2D + 2 D R (+ 2 D S) ] + sum_head 2 Cin Cout
  F     = 3 F_fwd                      (forward + dgrad + wgrad; recompute and the time-constant conditioning 1x1 not counted)
  B_alg = blocks (2 R e + 2 S' 4) + S' 4 + 8  +  blocks (3 R e + S' 4)      (block-fused ideal; e bytes / element,
                                                                            S' = skip width under use_skip else 0)
"""
from __future__ import annotations

import math
from typing import Dict, List


def dilations(kw: dict) -> List[List[int]]:
  """model.py:79-81,93-94: per-block dilation lists."""
  K = kw.get('kernel_size', 2)
  lpb, blocks = kw.get('layers_per_block', 1), kw.get('blocks', 10)
  max_power = int(math.log(kw.get('dilation_bound', 512), K))
  dil = [K ** (i % max_power) for i in range(lpb * blocks)]
  return [dil[b * lpb:(b + 1) * lpb] for b in range(blocks)]


def out_channels(kw: dict) -> int:
  return 3 * kw['num_mixtures'] if kw.get('num_mixtures') is not None else 2 ** kw.get('bits', 8)


def flops_fwd_per_sample(kw: dict) -> Dict[str, int]:
  K, R = kw.get('kernel_size', 2), kw.get('channels', 32)
  D = kw.get('dilation_channels') or R
  S = kw.get('skip_channels')
  dil = one = 0
  for dils in dilations(kw):
    cin = R
    for _ in range(len(dils) - 1):
      dil += 2 * K * cin * D
      cin = D
    dil += 2 * K * cin * 2 * D
    one += 2 * D * R + (2 * D * S if S is not None else 0)
  head = 0
  cin = (S if S is not None else R) if kw.get('use_skip', True) else R
  for ch in list(kw.get('final_layers_channels') or []) + [out_channels(kw)]:
    head += 2 * cin * ch
    cin = ch
  inp = 2 * K * R
  return {'dilated': dil, 'pointwise': one, 'head': head, 'input': inp, 'total': dil + one + head + inp}


def alg_bytes_per_sample(kw: dict, elem_bytes: int) -> int:
  """SURVEY.md 8(d) B_alg: HBM bytes per audio sample of a block-fused forward + backward pass."""
  R = kw.get('channels', 32)
  S = kw.get('skip_channels')
  blocks = kw.get('blocks', 10)
  sp = (S if S is not None else R) if kw.get('use_skip', True) else 0
  fwd = blocks * (2 * R * elem_bytes + 2 * sp * 4) + sp * 4 + 8
  bwd = blocks * (3 * R * elem_bytes + sp * 4)
  return fwd + bwd


def head_widths(kw: dict) -> List[int]:
  """[input width, hidden widths..., output channels] of the head (model.py:105-119)."""
  R = kw.get('channels', 32)
  S = kw.get('skip_channels')
  cin = (S if S is not None else R) if kw.get('use_skip', True) else R
  return [cin] + list(kw.get('final_layers_channels') or []) + [out_channels(kw)]


def hbm_phase_bytes_per_sample(kw: dict, elem_bytes: int, head_wgrads_grouped: bool = False) -> Dict[str, int]:
  """HBM bytes per audio sample each non-block phase of the step must move AS LAUNCHED (every launch reads its inputs
  and writes its outputs once; weights are negligible): the numerators of bench.py's `roofline_hbm` (north star item 3:
  skip accumulation, head and loss kernels with achieved HBM GB/s).  Keys = the phase labels of wn_profile_get.
  e = activation element size of the tier (2 bf16, 4 fp32); logits are always fp32."""
  R = kw.get('channels', 32)
  D = kw.get('dilation_channels') or R
  S = kw.get('skip_channels')
  L = kw.get('blocks', 10)
  e = elem_bytes
  C = out_channels(kw)
  ldd = (C + 63) // 64 * 64 if e == 2 else C
  hw = head_widths(kw)
  act = kw.get('activation') not in (None, 'linear')
  out = {
    'loss': 4 * C + 4 + e * ldd,                 # reads fp32 logits + the target sample, writes d logits
    'input_conv_fwd': 4 + e * R,                 # reads the audio sample (taps hit L1), writes h0
    'input_conv_bwd': 4 + e * R,                 # reads d h0 and the audio sample; the reduction partials are small
  }
  if kw.get('use_skip', True):
    sp = S if S is not None else R
    out['skip_sum'] = (L * D + sp) * e           # ONE K = L*D GEMM over the cached gate outputs (model.py:236)
  fwd = 0
  for i in range(len(hw) - 1):
    fwd += hw[i] * e + (4 * hw[i + 1] if i == len(hw) - 2 else hw[i + 1] * e)
  out['head_fwd'] = fwd
  bwd = 0
  for i in range(len(hw) - 2, -1, -1):
    gout = ldd if i == len(hw) - 2 else hw[i + 1]
    bwd += (gout + hw[i] + (hw[i] if (act and i > 0) else 0)) * e        # dgrad: reads dY (and the cached output), writes dX
    if not head_wgrads_grouped:
      bwd += (hw[i] + gout) * e                                          # its weight gradient: reads X and dY
  out['head_bwd'] = bwd
  return out
