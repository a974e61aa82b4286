"""Graph-replay time of the forward-only step (test_step: forward + loss) of a BASELINE config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenets_b200 import CONFIGS, WaveNet, model_kwargs, synth
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = dict(CONFIGS[name]); kw = model_kwargs(cfg)
B, T = cfg['batch_size'], cfg['recording_length']
cond_in = cfg.get('n_speakers', 109) if kw['conditioning'] == 'global' else 0
m = WaveNet(**kw, precision=cfg.get('precision', 'bf16'), max_batch=B, max_time=T)
m.build(((B, T, 1), (B, cond_in)) if cond_in else (B, T, 1))
x = torch.from_numpy(synth.frames(B, T, seed=0, apply_mulaw=cfg.get('apply_mulaw', True))).cuda()
c = torch.from_numpy(synth.speakers_onehot(B, cond_in, seed=0)).cuda() if cond_in else None
data = (x, c) if c is not None else x
for _ in range(4): m._step(data, False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): loss = m._step(data, False)
e1.record(); torch.cuda.synchronize()
print(f'{name} forward+loss: {e0.elapsed_time(e1) / steps:.3f} ms/step  loss {float(loss[0]):.4f}')
