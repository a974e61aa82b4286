"""Generate tests/golden/*.npz by running the REFERENCE'S OWN SOURCE (TEST INFRASTRUCTURE ONLY).

  python oracle/make_golden.py            # needs /root/reference (this container only)

/root/reference/src/layers.py and src/model.py are imported UNMODIFIED; `import tensorflow`
inside them resolves to oracle/tf_shim/tensorflow (a restatement of the few TF/Keras primitives
the reference calls, on torch CPU float64 — see that file's header for each definition).  So the
vectors below pin everything the reference's repository itself defines on the hot path:
constructor validation, dilation schedule and receptive field (model.py:79-81,122), layer wiring
and split order (layers.py:178-224), skip aliasing, conditioning broadcast (model.py:221-225),
shift-by-one slicing (model.py:319-321), the three losses (model.py:505-551), loss scaling and
L2 (model.py:328-334), what reaches `optimizer.apply_gradients` (model.py:335-336), the
deterministic branch of `sample_waveform` (model.py:393-503) and Keras' variable order.
What stays restated (not executed from upstream) are the TF primitives themselves.

The fixtures travel with the repo; nothing under tests/ reads /root/reference at run time.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('WN_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import tensorflow as tf  # noqa: E402  (the shim)
from src.layers import WaveNetLayer  # noqa: E402  (reference source)
from src.model import WaveNet  # noqa: E402  (reference source)

from oracle import wavenet_oracle as wo  # noqa: E402
from tests.util import oracle_config  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')

# name -> (model kwargs, cond_in, B, T, n_replicas)
CASES = {
  'cat_multidil': (dict(channels=16, blocks=3, layers_per_block=3, activation='leaky_relu', dilation_bound=16,
                        final_layers_channels=[32, 48], bits=8), 0, 2, 80, 1),
  'cond_skip': (dict(channels=16, blocks=4, layers_per_block=1, dilation_bound=8, final_layers_channels=[32],
                     skip_channels=24, dilation_channels=8, conditioning='global', mapping_layers=[8, 12],
                     mapping_activation='leaky_relu', activation='relu'), 7, 3, 61, 1),
  'logistic_cond': (dict(channels=16, blocks=3, layers_per_block=2, activation='tanh', dilation_bound=4,
                         final_layers_channels=[24], num_mixtures=5, sampling_function='logistic', bits=16,
                         conditioning='global', mapping_layers=[6], mapping_activation='relu'), 5, 2, 70, 1),
  'gaussian_noskip_k3': (dict(channels=8, blocks=3, layers_per_block=1, dilation_bound=9, final_layers_channels=[12],
                              num_mixtures=3, sampling_function='gaussian', use_skip=False, skip_channels=5,
                              kernel_size=3, use_residual=False), 0, 2, 50, 1),
  'l2_replicas2': (dict(channels=8, blocks=2, layers_per_block=2, dilation_bound=4, final_layers_channels=[8],
                        l2_reg_factor=0.01, activation='sigmoid', conditioning='global', mapping_layers=[4],
                        mapping_activation='tanh'), 5, 2, 40, 2),
  # sized for the bf16 tcgen05 tier (channel counts multiples of 64)
  'tc_cond_skip64': (dict(channels=64, blocks=3, layers_per_block=1, dilation_bound=8, final_layers_channels=[64],
                          skip_channels=64, conditioning='global', mapping_layers=[8, 16],
                          mapping_activation='leaky_relu', activation='leaky_relu'), 7, 2, 200, 1),
  'tc_multidil64': (dict(channels=64, blocks=2, layers_per_block=3, dilation_bound=8, final_layers_channels=[64],
                         activation='leaky_relu'), 0, 2, 150, 1),
  # dropout with an injected keep-mask (TF's RNG stream is not reproducible; the mask is part of the fixture)
  'dropout_mask': (dict(channels=16, blocks=2, layers_per_block=2, dilation_bound=4, final_layers_channels=[16],
                        activation='leaky_relu', dropout=0.25, skip_channels=16), 0, 2, 48, 1),
  # saturated softmax: the logits conv is scaled up (WEIGHT_SCALE) until most class probabilities and a part of the TARGET
  # probabilities leave [1e-7, 1 - 1e-7] -- Keras 3's clip inside sparse_categorical_crossentropy (model.py:516) is active
  'cat_saturated': (dict(channels=16, blocks=2, layers_per_block=1, dilation_bound=4, final_layers_channels=[32, 32],
                         activation='tanh', skip_channels=16, bits=8), 0, 2, 96, 1),
}
# per-case multipliers on the seeded weights (stored scaled in the fixture)
WEIGHT_SCALE = {'cat_saturated': {'final2/kernel': 150.0}}


def make_inputs(B, T, cond_in, seed):
  rng = np.random.default_rng(seed)
  x = np.clip(rng.standard_normal((B, T + 1, 1)) * 0.4, -1, 1).astype(np.float32)
  cond = np.eye(cond_in, dtype=np.float32)[rng.integers(0, cond_in, B)] if cond_in else None
  return x, cond


def ref_variables(model):
  """(oracle name, reference Variable) pairs, by walking the reference object graph."""
  out = [('causal/kernel', model.causal.kernel), ('causal/bias', model.causal.bias)]
  for b, blk in enumerate(model.wavenet_blocks):
    for j, conv in enumerate(blk.dilated_stack):
      out += [(f'block{b}/dil{j}/kernel', conv.kernel), (f'block{b}/dil{j}/bias', conv.bias)]
    out += [(f'block{b}/conv1/kernel', blk.conv1.kernel), (f'block{b}/conv1/bias', blk.conv1.bias)]
    if blk.conv_skip is not None:
      out += [(f'block{b}/conv_skip/kernel', blk.conv_skip.kernel), (f'block{b}/conv_skip/bias', blk.conv_skip.bias)]
    if blk.condition:
      out += [(f'block{b}/conv_cond/kernel', blk.conv_cond.kernel), (f'block{b}/conv_cond/bias', blk.conv_cond.bias)]
  for i, conv in enumerate(model.final):
    out += [(f'final{i}/kernel', conv.kernel), (f'final{i}/bias', conv.bias)]
  if model.conditioning == 'global':
    i = 0
    for layer in model.mapping.layers:
      if hasattr(layer, 'kernel'):
        out += [(f'mapping{i}/kernel', layer.kernel), (f'mapping{i}/bias', layer.bias)]
        i += 1
  return out


def run_case(name, kw, cond_in, B, T, nrep):
  tf.set_num_replicas(nrep)
  cfg = oracle_config(kw, cond_in)
  x, cond = make_inputs(B, T, cond_in, seed=abs(hash(name)) % 1000 if False else sum(map(ord, name)))
  model = WaveNet(**kw)
  opt = tf.RecordingOptimizer()
  model.compile(optimizer=opt)
  xt = torch.tensor(x, dtype=torch.float64)
  ct = torch.tensor(cond, dtype=torch.float64) if cond is not None else None
  inputs = [xt[:, :-1], ct] if ct is not None else xt[:, :-1]
  model(inputs)                                   # build-by-call, like train.py:232-235
  # ---- weights: oracle-named, seeded, rounded to fp32 so every tier sees identical values
  p = {k: v.astype(np.float32) for k, v in wo.init_params(cfg, seed=1).items()}
  for k, f in WEIGHT_SCALE.get(name, {}).items():
    p[k] = (p[k] * np.float32(f)).astype(np.float32)
  pairs = ref_variables(model)
  assert [n for n, _ in pairs] == [n for n, _ in wo.param_specs(cfg)], 'variable naming / order'
  assert all(a is b for (_, a), b in zip(pairs, model.trainable_variables)), 'Keras tracking order'
  for n, v in pairs:
    assert tuple(v.shape) == p[n].shape, (n, v.shape, p[n].shape)
    v.assign(p[n].astype(np.float64))
  out = {'x': x, 'n_replicas': np.int64(nrep), 'B': np.int64(B), 'T': np.int64(T), 'cond_in': np.int64(cond_in)}
  if cond is not None:
    out['cond'] = cond
  for n, w in p.items():
    out['w/' + n] = w
  # ---- dropout: inject a seeded keep-mask per block (applies to the block input, layers.py:195-196)
  if kw.get('dropout', 0) > 0:
    rng = np.random.default_rng(7)
    for b, blk in enumerate(model.wavenet_blocks):
      keep = rng.random((B, T, kw['channels'])) >= kw['dropout']
      blk.dropout.mask = torch.tensor(keep)
      out[f'keep/block{b}'] = np.packbits(keep.reshape(-1))
  # ---- structure
  out['receptive_field'] = np.int64(model.receptive_field)
  out['dilations'] = np.array([[c.dilation_rate for c in blk.dilated_stack] for blk in model.wavenet_blocks], dtype=np.int64)
  # ---- forward (training=False): probabilities / mixture parameters
  with torch.no_grad():
    pred = model(inputs, training=False)
    # kept at a subset of time steps (first 4: causal edge; last 16: deepest context; every 8th in between)
    tsel = np.unique(np.concatenate([np.arange(min(4, T)), np.arange(0, T, 8), np.arange(max(0, T - 16), T)]))
    out['pred_t'] = tsel.astype(np.int64)
    out['pred'] = pred.numpy()[:, tsel, :].astype(np.float32)
    y = xt[:, 1:, :]
    tgt = model.prepare_target(y)
    out['target'] = tgt.numpy()
    out['loss_per_sample'] = model.loss_fn(tgt, pred).numpy().reshape(B, T).astype(np.float64)
    out['sample_deterministic'] = model.sample_waveform(pred, deterministic=True).numpy().astype(np.float32)
  # ---- test_step and train_step through the reference's own methods
  data = (xt, ct) if ct is not None else xt
  with torch.no_grad():
    out['test_loss'] = np.float64(model.test_step(data)['loss'])
  model.compile(optimizer=opt)                    # fresh metric trackers
  res = model.train_step(data)
  out['train_loss'] = np.float64(res['loss'])
  if 'reg_loss' in res:
    out['reg_loss'] = np.float64(res['reg_loss'])
  by_id = {id(v): g for g, v in zip(opt.gradients, opt.variables)}
  for n, v in pairs:
    out['g/' + n] = by_id[id(v)].numpy().astype(np.float32)
  # ---- one block on its own (WaveNetLayer.call): block 0 on the causal conv's output
  with torch.no_grad():
    h0 = model.causal(inputs[0] if ct is not None else inputs)
    blk = model.wavenet_blocks[0]
    if ct is not None:
      c = model.mapping(ct)
      c = tf.repeat(tf.expand_dims(c, axis=1), h0.shape[1], axis=1)
      xo, sk = blk([h0, c], training=False)
      out['layer0/cond'] = c[:, 0, :].numpy().astype(np.float32)
    else:
      xo, sk = blk(h0, training=False)
    out['layer0/x'] = h0.numpy().astype(np.float32)
    out['layer0/x_out'] = xo.numpy().astype(np.float32)
    out['layer0/skip'] = sk.numpy().astype(np.float32)
    if kw.get('dropout', 0) > 0 and ct is None:
      # WaveNetLayer.call(training=True): dropout on the conv branch with the injected keep-mask of block 0 (layers.py:192-196)
      xo, sk = blk(h0, training=True)
      out['layer0/x_out_train'] = xo.numpy().astype(np.float32)
      out['layer0/skip_train'] = sk.numpy().astype(np.float32)
    if cfg.num_mixtures is None:
      pn = pred.numpy()
      out['clip_stats'] = np.array([((pn < 1e-7) | (pn > 1 - 1e-7)).mean(),
                                    (np.take_along_axis(pn, tgt.numpy().reshape(B, T, 1).astype(np.int64), -1) < 1e-7).mean()])
  os.makedirs(OUT, exist_ok=True)
  np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
  nbytes = os.path.getsize(os.path.join(OUT, name + '.npz'))
  if 'clip_stats' in out:
    print(f'{name:22s} clipped probabilities {out["clip_stats"][0]:.3f}, clipped targets {out["clip_stats"][1]:.3f}')
  print(f'{name:22s} loss={out["train_loss"]:.6f} test={out["test_loss"]:.6f} rf={int(out["receptive_field"])} '
        f'params={sum(w.size for w in p.values())} file={nbytes / 1024:.0f} KB')


def known_answers():
  """Reference-derived scalars: receptive fields / dilation schedules of the BASELINE configs, quantiser
  edges (model.py:151-153), constructor validation messages (model.py:52-70)."""
  out = {}
  defaults = dict(kernel_size=2, channels=32, blocks=5, layers_per_block=5, dilation_bound=256, activation='leaky_relu',
                  final_layers_channels=[128, 256])
  m = WaveNet(**defaults)
  out['defaults_rf'] = np.int64(m.receptive_field)
  out['defaults_dilations'] = np.array([[c.dilation_rate for c in b.dilated_stack] for b in m.wavenet_blocks], dtype=np.int64)
  m = WaveNet(kernel_size=2, channels=32, blocks=40, layers_per_block=1, dilation_bound=1024, final_layers_channels=[], use_skip=False)
  out['c5_rf'] = np.int64(m.receptive_field)
  m = WaveNet(kernel_size=2, channels=32, blocks=30, layers_per_block=1, dilation_bound=1024, final_layers_channels=[])
  out['c2_rf'] = np.int64(m.receptive_field)
  for bits in (8, 16):
    m = WaveNet(channels=8, blocks=1, final_layers_channels=[], dilation_bound=2, bits=bits)
    edges = np.linspace(-1, 1, 2 ** bits + 1).astype(np.float32)
    x = np.concatenate([edges, np.nextafter(edges, np.float32(2)), np.nextafter(edges, np.float32(-2)),
                        np.array([-1e-30, 1e-30, -0.0, 0.0, 1.0, -1.0, 0.99999994, -0.99999994], np.float32),
                        np.random.default_rng(0).uniform(-1, 1, 4096).astype(np.float32)])
    out[f'quant{bits}_x'] = x
    out[f'quant{bits}_idx'] = m.prepare_target(torch.tensor(x, dtype=torch.float64)).numpy().astype(np.int64)
  msgs = []
  bad = [dict(conditioning='x'), dict(kernel_size=1), dict(dilation_bound=100), dict(layers_per_block=0), dict(blocks=0),
         dict(num_mixtures=0, sampling_function='logistic'), dict(dropout=1.5), dict(sampling_function='foo'),
         dict(sampling_function='categorical', num_mixtures=3)]
  for kw in bad:
    try:
      WaveNet(**{**dict(final_layers_channels=[]), **kw})
      msgs.append('')
    except ValueError as e:
      msgs.append(str(e))
  out['value_errors'] = np.array(msgs)
  np.savez_compressed(os.path.join(OUT, 'known_answers.npz'), **out)
  print('known_answers: defaults rf', int(out['defaults_rf']), 'c2 rf', int(out['c2_rf']), 'c5 rf', int(out['c5_rf']))


if __name__ == '__main__':
  torch.manual_seed(0)
  only = sys.argv[1:]          # `python oracle/make_golden.py [case ...]`: regenerate only the named fixtures
  for name, (kw, cond_in, B, T, nrep) in CASES.items():
    if not only or name in only:
      run_case(name, kw, cond_in, B, T, nrep)
  tf.set_num_replicas(1)
  if not only or 'known_answers' in only:
    known_answers()
