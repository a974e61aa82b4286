"""wavenets_b200 — B200-native (sm_100a) WaveNet residual-stack training pass.

Drop-in for the hot path of jirsat/wavenets (`src/layers.py`, `src/model.py`): the same
`WaveNetLayer` / `WaveNet` Python signatures over hand-written CUDA kernels behind a C ABI
(include/wavenet_b200.h, libwavenet_b200.so).  No CPU fallback.
"""
from .layers import WaveNetLayer
from .model import WaveNet
from .config import load_config, model_kwargs, CONFIGS
from . import synth
from . import optimizers

__all__ = ['WaveNetLayer', 'WaveNet', 'load_config', 'model_kwargs', 'CONFIGS', 'synth', 'optimizers']
