"""In-situ per-launch durations (CUDA events around every launch, eager single-stream mode, warm L2) of one C2 step."""
import collections, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenets_b200 import CONFIGS, WaveNet, model_kwargs, synth
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
cfg = dict(CONFIGS[name]); kw = model_kwargs(cfg)
B, T = cfg['batch_size'], cfg['recording_length']
if len(sys.argv) > 2: B = int(sys.argv[2])
cond_in = cfg.get('n_speakers', 109) if kw['conditioning'] == 'global' else 0
m = WaveNet(**kw, precision=cfg.get('precision', 'bf16'), max_batch=B, max_time=T)
m.build(((B, T, 1), (B, cond_in)) if cond_in else (B, T, 1))
x = torch.from_numpy(synth.frames(B, T, seed=0, apply_mulaw=cfg.get('apply_mulaw', True))).cuda()
c = torch.from_numpy(synth.speakers_onehot(B, cond_in, seed=0)).cuda() if cond_in else None
data = (x, c) if c is not None else x
h = m.handle
for _ in range(3): m.train_step_async(data)
torch.cuda.synchronize()
h.lib.wn_profile_begin(h.h, 4)
for _ in range(3): m.train_step_async(data)
ms = C.c_double(); n = C.c_int64()
h.lib.wn_profile_end(h.h, C.byref(ms), C.byref(n))
cnt = h.lib.wn_profile_get(h.h, 0, None, None, 0)
agg = collections.defaultdict(lambda: [0, 0.0])
lab = C.create_string_buffer(64); d = C.c_double()
for i in range(cnt):
  h.lib.wn_profile_get(h.h, i, C.byref(d), lab, 64)
  a = agg[lab.value.decode()]; a[0] += 1; a[1] += d.value
tot = sum(v[1] for v in agg.values())
print(f'{cnt // 3} launches/step, {tot / 3:.3f} ms/step summed (serial, warm)')
for k, (nn, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print(f'{t / 3:8.3f} ms/step {nn // 3:4d}x {1e3 * t / nn:8.1f} us  {k}')
if '--misc' in sys.argv:
  per = cnt // 3
  for i in range(2 * per, cnt):
    h.lib.wn_profile_get(h.h, i, C.byref(d), lab, 64)
    if lab.value.decode() == 'misc':
      print(f'  launch {i - 2 * per:4d}  {1e3 * d.value:8.1f} us')
