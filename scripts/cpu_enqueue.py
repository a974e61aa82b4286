"""Host-side enqueue cost of one training step vs its GPU duration."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenets_b200 import CONFIGS, WaveNet, model_kwargs, synth
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
cfg = dict(CONFIGS[name]); kw = model_kwargs(cfg)
B, T = cfg['batch_size'], cfg['recording_length']
cond_in = cfg.get('n_speakers', 109) if kw['conditioning'] == 'global' else 0
m = WaveNet(**kw, precision=cfg.get('precision', 'bf16'), max_batch=B, max_time=T)
m.build(((B, T, 1), (B, cond_in)) if cond_in else (B, T, 1))
x = torch.from_numpy(synth.frames(B, T, seed=0)).cuda()
c = torch.from_numpy(synth.speakers_onehot(B, cond_in, seed=0)).cuda() if cond_in else None
d = (x, c) if c is not None else x
for _ in range(3): m.train_step_async(d)
torch.cuda.synchronize()
for it in range(3):
  t0 = time.perf_counter()
  m.train_step_async(d)
  t1 = time.perf_counter()
  torch.cuda.synchronize()
  t2 = time.perf_counter()
  print(f'enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms')
