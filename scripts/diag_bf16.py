"""Diagnostic: per-tensor bf16-tier error vs the fp64 oracle (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import wavenet_oracle as wo
from tests.util import rel_l2
from tests.test_gpu_parity_bf16 import BF16_MODELS, _build

names = sys.argv[1:] or sorted(BF16_MODELS)
for name in names:
  for B, T in [(2, 100), (3, 333)]:
    m, cfg, p, x, cond = _build(BF16_MODELS[name], B, T)
    c64 = None if cond is None else cond.astype(np.float64)
    loss_o, g_o, _ = wo.train_step(p, cfg, x.astype(np.float64), c64)
    out = m.train_step((x, cond) if cond is not None else x)
    g = m.get_grads()
    pred_o, _ = wo.model_forward(p, cfg, x[:, :-1].astype(np.float64), c64)
    pred = m((x[:, :-1], cond) if cond is not None else x[:, :-1]).cpu().numpy()
    errs = sorted(((rel_l2(g[k], g_o[k]), k) for k in g_o if np.linalg.norm(g_o[k]) > 0), reverse=True)
    print(f'{name} B={B} T={T}: loss {out["loss"]:.5f} vs {loss_o:.5f} (rel {abs(out["loss"]-loss_o)/abs(loss_o):.2e}) pred relL2 {rel_l2(pred, pred_o):.2e}')
    print('   worst grads:', ', '.join(f'{k}={e:.2e}' for e, k in errs[:6]), ' median', f'{np.median([e for e,_ in errs]):.2e}')
