"""Weights-only checkpoints and resume — stands where the reference's `ModelCheckpoint(save_weights_only=True)` /
`load_weights` / resume-by-filename logic does (train.py:68-86,149-154,237-238).

Format: `.npz` keyed by the Keras variable names of `WaveNet.variable_names`, Keras layouts (conv kernel (K,Cin,Cout),
bias (Cout,), dense kernel (in,out)); h5py is not required.  File names keep the reference's pattern
`weights-e{epoch:04d}-lr{lr}.weights.<ext>`, from which `find_last_checkpoint` recovers the epoch and the learning
rate exactly like train.py:76-86 (lexicographically last file; optimizer slots are not saved, as upstream)."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def checkpoint_name(epoch: int, lr: float, ext: str = 'npz') -> str:
  return f'weights-e{epoch:04d}-lr{lr}.weights.{ext}'


def save_weights(model, path: str) -> str:
  w = model.get_weights()
  if any('__' in k for k in w):
    raise ValueError('variable names containing "__" cannot be stored ("/" is written as "__" in the .npz keys)')
  os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
  with open(path, 'wb') as f:
    np.savez(f, **{k.replace('/', '__'): v for k, v in w.items()})
  return path


def load_weights(model, path: str, strict: bool = True) -> None:
  if path.endswith(('.h5', '.hdf5', '.keras')):
    # the reference's ModelCheckpoint writes Keras `.weights.h5` files (train.py:149-154); h5py is not a dependency here
    raise ValueError(f'{path}: Keras HDF5 checkpoints are not read here (export the variables in trainable_variables order to .npz: '
                     'np.savez(path, **{name.replace("/", "__"): array}) with the names of WaveNet.variable_names)')
  z = np.load(path)
  w = {k.replace('__', '/'): z[k] for k in z.files}
  if model.built:
    names = set(model.variable_names)
    missing, extra = names - set(w), set(w) - names
    if strict and (missing or extra):
      raise ValueError(f'checkpoint does not match the model: missing {sorted(missing)[:4]}, unexpected {sorted(extra)[:4]}')
    w = {k: v for k, v in w.items() if k in names}
  model.set_weights(w)


def find_last_checkpoint(results_dir: str) -> Optional[Tuple[str, int, float]]:
  """train.py:68-86: sort the directory listing, take the last entry, parse `-e{epoch}` and `-lr{lr}` out of its name."""
  try:
    checkpoints = sorted(os.listdir(results_dir))
  except FileNotFoundError:
    return None
  # (only the files this module can read: a directory shared with a reference run also holds `.weights.h5` files)
  checkpoints = [c for c in checkpoints if '.weights' in c and c.endswith('.npz')]
  if not checkpoints:
    return None
  name = checkpoints[-1]
  filename = name.split('.weights')[0]
  filename, learning_rate = filename.split('-lr')
  return os.path.join(results_dir, name), int(filename.split('-e')[-1]), float(learning_rate)
