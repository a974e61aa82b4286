"""Config handling compatible with the reference YAML files (train.py:22-60).

Key names are the reference's, including the `use_resiudal` typo (train.py:46,221;
configfiles/defaults.yaml:24), so reference YAMLs load unchanged.
"""
from __future__ import annotations

import copy

# train.py:22-50
DEFAULTS = {
  'epochs': 500, 'lr': 0.0005, 'recording_length': 8000, 'batch_size': 64, 'apply_mulaw': False,
  'jit_compile': False, 'dataset': './datasets/vctk8000',
  'kernel_size': 2, 'channels': 32, 'blocks': 5, 'layers_per_block': 5, 'activation': 'leaky_relu',
  'conditioning': 'global', 'mapping_layers': [8, 16, 32], 'mapping_activation': 'leaky_relu',
  'dropout': 0.1, 'dilation_bound': 256, 'num_mixtures': 8, 'sampling_function': 'gaussian', 'bits': 16,
  'skip_channels': None, 'dilation_channels': None, 'use_resiudal': True, 'use_skip': True,
  'final_layers_channels': [128, 256], 'l2_reg_factor': 0,
}


def load_config(path=None) -> dict:
  cfg = copy.deepcopy(DEFAULTS)
  if path is not None:
    import yaml
    with open(path) as f:
      cfg.update(yaml.safe_load(f) or {})
  return cfg


def model_kwargs(cfg: dict) -> dict:
  """The WaveNet(...) call of train.py:206-224."""
  return dict(
    kernel_size=cfg['kernel_size'], channels=cfg['channels'], blocks=cfg['blocks'],
    layers_per_block=cfg['layers_per_block'], activation=cfg['activation'], conditioning=cfg['conditioning'],
    mapping_layers=cfg['mapping_layers'], mapping_activation=cfg['mapping_activation'], dropout=cfg['dropout'],
    dilation_bound=cfg['dilation_bound'], num_mixtures=cfg['num_mixtures'],
    sampling_function=cfg['sampling_function'], bits=cfg['bits'], skip_channels=cfg['skip_channels'],
    dilation_channels=cfg['dilation_channels'], use_residual=cfg['use_resiudal'], use_skip=cfg['use_skip'],
    final_layers_channels=cfg['final_layers_channels'], l2_reg_factor=cfg['l2_reg_factor'])


def _c(**over):
  c = copy.deepcopy(DEFAULTS)
  c.update(dropout=0.0)
  c.update(over)
  return c


# The five BASELINE.json configurations (widths for C2-C5 are SURVEY.md 8d's proposal).
CONFIGS = {
  # C1: defaults.yaml topology, 256-way mu-law softmax, no conditioning, batch 1, T=8000
  'c1': _c(conditioning=None, num_mixtures=None, sampling_function='categorical', bits=8, apply_mulaw=True,
           batch_size=1, precision='fp32'),
  # C2: global conditioning (109 speakers), 30x1, R=D=S=256, softmax-256, bf16, 8 sequences / GPU
  'c2': _c(channels=256, skip_channels=256, blocks=30, layers_per_block=1, dilation_bound=1024,
           final_layers_channels=[256, 256], num_mixtures=None, sampling_function='categorical', bits=8,
           apply_mulaw=True, batch_size=8, precision='bf16', n_speakers=109),
  # C3: multiple dilated convs per layer (5x5 defaults schedule), R=D=256
  'c3': _c(channels=256, conditioning=None, num_mixtures=None, sampling_function='categorical', bits=8,
           apply_mulaw=True, batch_size=8, precision='bf16'),
  # C4: C2 with a 10-component mixture-of-logistics head
  'c4': _c(channels=256, skip_channels=256, blocks=30, layers_per_block=1, dilation_bound=1024,
           final_layers_channels=[256, 256], num_mixtures=10, sampling_function='logistic', bits=16,
           apply_mulaw=False, batch_size=8, precision='bf16', n_speakers=109),
  # C5: deep stack without skip connections, 4 x (1..512), T=16384
  'c5': _c(channels=256, blocks=40, layers_per_block=1, dilation_bound=1024, use_skip=False, conditioning=None,
           final_layers_channels=[256, 256], num_mixtures=None, sampling_function='categorical', bits=8,
           apply_mulaw=True, recording_length=16384, batch_size=4, precision='bf16'),
}
