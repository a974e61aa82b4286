#!/usr/bin/env python
"""train.py of the reference, end to end on synthetic audio, with every stage on the GPU through libwavenet_b200.so:

  config (reference YAML keys, train.py:22-60)  ->  device input pipeline (utils.py:22-85)  ->  WaveNet(...) (train.py:206-224)
  ->  compile(Adam(lr, clipnorm=1.0), metrics=[MSE]) (train.py:225-228)  ->  train_step loop (model.py:309-348)
  ->  weights-only checkpoints named like the reference's (train.py:149-154) and resume-by-filename (train.py:68-86)
  ->  generate (train.py:253-261).

  python examples/train_synthetic.py [--configfile some.yaml] [--steps 50] [--results ./results/demo]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavenets_b200 import WaveNet, checkpoint, data, load_config, model_kwargs  # noqa: E402
from wavenets_b200.metrics import MeanSquaredError  # noqa: E402
from wavenets_b200.optimizers import Adam  # noqa: E402


def synthetic_recordings(n, seconds, fs, seed=0):
  """int16 'recordings' (a few harmonics + noise), standing in for the VCTK clips the reference loads (train.py:90-126)."""
  rng = np.random.default_rng(seed)
  t = np.arange(int(seconds * fs)) / fs
  for i in range(n):
    f0 = rng.uniform(90, 300)
    x = sum(rng.uniform(0.05, 0.3) * np.sin(2 * np.pi * f0 * (k + 1) * t + rng.uniform(0, 6.28)) for k in range(4))
    x = x + rng.normal(0, 0.01, t.shape)
    yield (np.clip(x, -1, 1) * 32767).astype(np.int16), int(rng.integers(0, 2))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--configfile', default=None)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--results', default='./results/demo')
  ap.add_argument('--precision', default='bf16')
  args = ap.parse_args()
  cfg = load_config(args.configfile)
  if args.configfile is None:      # a small model that trains in seconds
    cfg.update(channels=64, skip_channels=64, blocks=6, layers_per_block=1, dilation_bound=64, final_layers_channels=[64],
               conditioning='global', mapping_layers=[8, 16], num_mixtures=None, sampling_function='categorical', bits=8,
               apply_mulaw=True, recording_length=2000, batch_size=8, dropout=0.05, lr=1e-3)
  T, B = cfg['recording_length'], cfg['batch_size']

  # ---- dataset: recordings -> frames of T+1 samples (hop T), mu-law, range filter, one-hot condition; all on the device
  frames, conds = [], []
  for speech, gender in synthetic_recordings(8, seconds=4.0, fs=8000):
    f, c = data.preprocess_recording(speech, T, cfg['apply_mulaw'], condition_id=gender if cfg['conditioning'] else None, condition_depth=2)
    frames.append(f)
    conds.append(c)
  frames = torch.cat(frames)
  conds = torch.cat(conds) if cfg['conditioning'] else None
  print(f'{frames.shape[0]} frames of {T + 1} samples, range [{float(frames.min()):.3f}, {float(frames.max()):.3f}]')

  # ---- resume (train.py:68-86): last checkpoint of the run directory gives the epoch and the learning rate
  initial_epoch, last = 0, checkpoint.find_last_checkpoint(args.results)
  if last is not None:
    path, initial_epoch, cfg['lr'] = last
    print(f'resuming from {path}: epoch {initial_epoch}, lr {cfg["lr"]}')

  model = WaveNet(**model_kwargs(cfg), precision=args.precision, max_batch=B, max_time=T)
  model.compile(optimizer=Adam(learning_rate=cfg['lr'], clipnorm=1.0), metrics=[MeanSquaredError()])
  model.build(((B, T, 1), (B, 2)) if cfg['conditioning'] else (B, T, 1))
  if last is not None:
    checkpoint.load_weights(model, last[0])
  print(f'receptive field {model.receptive_field} samples, {model.handle.n_scalars} parameters')

  rng = np.random.default_rng(1)
  t0, best, pending = time.perf_counter(), float('inf'), None
  steps_per_epoch = 20                       # Keras `fit` resets the metric trackers every epoch; the logs are running means
  for step in range(args.steps):
    if step and step % steps_per_epoch == 0:
      if pending is not None:
        pending.result()
      model.reset_metrics()
    idx = torch.from_numpy(rng.choice(frames.shape[0], B, replace=frames.shape[0] < B)).to(frames.device)
    batch = (frames[idx], conds[idx]) if conds is not None else frames[idx]
    nxt = model.train_step_deferred(batch)
    if pending is not None:
      logs = pending.result()
      if step % 10 == 0:
        print(f'step {step:4d}  ' + '  '.join(f'{k} {v:.4f}' for k, v in logs.items()))
      best = min(best, logs['loss'])
    pending = nxt
  logs = pending.result()
  logs = dict(model.last_step_logs)          # the last step's own values (the dict above holds the epoch's running means)
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  print(f'{args.steps} steps in {dt:.2f} s ({args.steps * B * T / dt / 1e6:.2f} M samples/s incl. optimizer and metrics), last loss {logs["loss"]:.4f}')
  path = checkpoint.save_weights(model, os.path.join(args.results, checkpoint.checkpoint_name(initial_epoch + 1, cfg['lr'])))
  print('saved', path)

  # ---- generation (train.py:253-261)
  cond = conds[:2] if conds is not None else None
  t0 = time.perf_counter()
  wav = model.generate(1000, condition=cond, batch_size=2, deterministic=False, seed=3)
  torch.cuda.synchronize()
  print(f'generated {tuple(wav.shape)} in {time.perf_counter() - t0:.2f} s; audio range after inverse mu-law '
        f'[{float(data.inverse_mu_law(wav).min()):.3f}, {float(data.inverse_mu_law(wav).max()):.3f}]')
  return logs


if __name__ == '__main__':
  main()
