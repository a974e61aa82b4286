"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the reference's own source)."""
import os
from types import SimpleNamespace

import numpy as np

from tests.util import oracle_config

GOLD = os.path.join(os.path.dirname(__file__), 'golden')

# must mirror oracle/make_golden.py:CASES (the kwargs are not stored in the .npz)
CASES = {
  'cat_multidil': (dict(channels=16, blocks=3, layers_per_block=3, activation='leaky_relu', dilation_bound=16,
                        final_layers_channels=[32, 48], bits=8), 0),
  'cond_skip': (dict(channels=16, blocks=4, layers_per_block=1, dilation_bound=8, final_layers_channels=[32],
                     skip_channels=24, dilation_channels=8, conditioning='global', mapping_layers=[8, 12],
                     mapping_activation='leaky_relu', activation='relu'), 7),
  'logistic_cond': (dict(channels=16, blocks=3, layers_per_block=2, activation='tanh', dilation_bound=4,
                         final_layers_channels=[24], num_mixtures=5, sampling_function='logistic', bits=16,
                         conditioning='global', mapping_layers=[6], mapping_activation='relu'), 5),
  'gaussian_noskip_k3': (dict(channels=8, blocks=3, layers_per_block=1, dilation_bound=9, final_layers_channels=[12],
                              num_mixtures=3, sampling_function='gaussian', use_skip=False, skip_channels=5,
                              kernel_size=3, use_residual=False), 0),
  'l2_replicas2': (dict(channels=8, blocks=2, layers_per_block=2, dilation_bound=4, final_layers_channels=[8],
                        l2_reg_factor=0.01, activation='sigmoid', conditioning='global', mapping_layers=[4],
                        mapping_activation='tanh'), 5),
  'tc_cond_skip64': (dict(channels=64, blocks=3, layers_per_block=1, dilation_bound=8, final_layers_channels=[64],
                          skip_channels=64, conditioning='global', mapping_layers=[8, 16],
                          mapping_activation='leaky_relu', activation='leaky_relu'), 7),
  'tc_multidil64': (dict(channels=64, blocks=2, layers_per_block=3, dilation_bound=8, final_layers_channels=[64],
                         activation='leaky_relu'), 0),
  'dropout_mask': (dict(channels=16, blocks=2, layers_per_block=2, dilation_bound=4, final_layers_channels=[16],
                        activation='leaky_relu', dropout=0.25, skip_channels=16), 0),
  # logits conv scaled x150 in the fixture: 39 % of the class probabilities and 38 % of the target probabilities are
  # outside [1e-7, 1 - 1e-7], so Keras 3's clip inside sparse_categorical_crossentropy (model.py:516) is active
  'cat_saturated': (dict(channels=16, blocks=2, layers_per_block=1, dilation_bound=4, final_layers_channels=[32, 32],
                         activation='tanh', skip_channels=16, bits=8), 0),
}


def load_case(name):
  kw, cond_in = CASES[name]
  z = np.load(os.path.join(GOLD, name + '.npz'))
  B, T = int(z['B']), int(z['T'])
  assert int(z['cond_in']) == cond_in
  c = SimpleNamespace(name=name, kw=kw, cond_in=cond_in, cfg=oracle_config(kw, cond_in), B=B, T=T,
                      n_replicas=int(z['n_replicas']), x=z['x'], cond=z['cond'] if 'cond' in z.files else None)
  c.weights = {k[2:]: z[k] for k in z.files if k.startswith('w/')}
  c.grads = {k[2:]: z[k].astype(np.float64) for k in z.files if k.startswith('g/')}
  c.receptive_field, c.dilations = int(z['receptive_field']), z['dilations']
  c.pred_t, c.pred, c.target = z['pred_t'], z['pred'].astype(np.float64), z['target']
  c.loss_per_sample, c.sample_deterministic = z['loss_per_sample'], z['sample_deterministic'].astype(np.float64)
  c.test_loss, c.train_loss = float(z['test_loss']), float(z['train_loss'])
  c.reg_loss = float(z['reg_loss']) if 'reg_loss' in z.files else None
  c.layer0_x, c.layer0_x_out, c.layer0_skip = z['layer0/x'], z['layer0/x_out'].astype(np.float64), z['layer0/skip'].astype(np.float64)
  c.layer0_cond = z['layer0/cond'] if 'layer0/cond' in z.files else None
  c.layer0_x_out_train = z['layer0/x_out_train'].astype(np.float64) if 'layer0/x_out_train' in z.files else None
  c.layer0_skip_train = z['layer0/skip_train'].astype(np.float64) if 'layer0/skip_train' in z.files else None
  c.clip_stats = z['clip_stats'] if 'clip_stats' in z.files else None
  c.keep_masks = None
  if kw.get('dropout', 0) > 0:
    n = B * T * kw['channels']
    c.keep_masks = [np.unpackbits(z[f'keep/block{b}'])[:n].reshape(B, T, kw['channels']).astype(bool) for b in range(kw['blocks'])]
  return c
