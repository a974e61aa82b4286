"""Generation speed (samples/s) of the per-conv-history step form on a BASELINE config."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenets_b200 import CONFIGS, WaveNet, model_kwargs, synth
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
length = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
cfg = dict(CONFIGS[name]); kw = model_kwargs(cfg)
cond_in = cfg.get('n_speakers', 109) if kw['conditioning'] == 'global' else 0
m = WaveNet(**kw, precision=cfg.get('precision', 'bf16'), max_batch=B, max_time=256)
m.build(((B, 256, 1), (B, cond_in)) if cond_in else (B, 256, 1))
c = torch.from_numpy(synth.speakers_onehot(B, cond_in, seed=0)).cuda() if cond_in else None
prime = torch.from_numpy(synth.frames(B, m.receptive_field - 1, seed=0)).cuda()
m.generate(8, condition=c, sample=prime, deterministic=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = m.generate(length, condition=c, sample=prime, deterministic=False, seed=1)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
steps = m.receptive_field - 1 + length
print(f'{name} B={B}: primed {m.receptive_field} + generated {length} samples/sequence in {dt:.2f} s -> {steps / dt:.0f} steps/s, '
      f'{B * steps / dt:.0f} samples/s over the batch; out range [{float(out.min()):.3f}, {float(out.max()):.3f}]')
