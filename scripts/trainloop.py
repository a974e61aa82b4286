"""Full training iteration (fwd + loss + bwd + clipnorm + Adam + re-pack + sampled-waveform MSE) vs the fwd+bwd step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenets_b200 import CONFIGS, WaveNet, model_kwargs, synth
from wavenets_b200.optimizers import Adam
from wavenets_b200.metrics import MeanSquaredError
cfg = dict(CONFIGS['c2']); kw = model_kwargs(cfg)
B, T = cfg['batch_size'], cfg['recording_length']
m = WaveNet(**kw, precision='bf16', max_batch=B, max_time=T)
m.build(((B, T, 1), (B, 109)))
x = torch.from_numpy(synth.frames(B, T, seed=0)).cuda()
c = torch.from_numpy(synth.speakers_onehot(B, 109, seed=0)).cuda()
def run(n):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(n): out = m.train_step((x, c))
  torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, out
for _ in range(4): m.train_step((x, c))
ms0, out0 = run(20)
m.compile(optimizer=Adam(learning_rate=5e-4, clipnorm=1.0), metrics=[MeanSquaredError()])
for _ in range(3): m.train_step((x, c))
ms1, out1 = run(20)
print(f'fwd+bwd step (sync per step): {ms0:.3f} ms   full iteration with Adam+clipnorm+repack+MSE: {ms1:.3f} ms   {out1}')
