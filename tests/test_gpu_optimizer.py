"""`tf.keras.optimizers.Adam(lr, clipnorm=1.0)` + `optimizer.apply_gradients` (train.py:225-226,
model.py:336) on the GPU vs the oracle restatement of Keras 3 Adam, over several training steps."""
import numpy as np
import pytest

from oracle import wavenet_oracle as wo
from tests.util import COND_IN, SMALL_MODELS, make_inputs, oracle_config, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,precision,tol', [('cond_skip', 'fp32', 1e-4), ('categorical_multidil', 'fp32', 1e-4), ('l2', 'fp32', 1e-4)])
def test_adam_clipnorm_trajectory(name, precision, tol):
  from wavenets_b200 import WaveNet
  from wavenets_b200.optimizers import Adam
  kw = SMALL_MODELS[name]
  cond_in = COND_IN if kw.get('conditioning') else 0
  cfg = oracle_config(kw, cond_in)
  p = {k: v.astype(np.float32).astype(np.float64) for k, v in wo.init_params(cfg, seed=1).items()}
  x, cond = make_inputs(3, 97, cond_in)
  m = WaveNet(**kw, precision=precision)
  opt = Adam(learning_rate=3e-3, clipnorm=1.0)
  m.compile(optimizer=opt)
  m.build((x[:, :-1].shape, cond.shape) if cond is not None else x[:, :-1].shape)
  m.set_weights({k: v.astype(np.float32) for k, v in p.items()})
  data = (x, cond) if cond is not None else x
  state = {}
  c64 = None if cond is None else cond.astype(np.float64)
  clipped_any = False
  for step in range(4):
    loss_o, g_o, aux = wo.train_step(p, cfg, x.astype(np.float64), c64)
    m.train_step(data)
    out = m.last_step_logs          # (a compiled model's train_step returns Keras' running means, model.py:340-348)
    assert abs(out['loss'] - aux['loss_no_reg']) <= 10 * tol * abs(loss_o), (step, out['loss'], aux['loss_no_reg'])
    norms = opt.grad_norms(m)
    for k, g in g_o.items():
      n_o = float(np.sqrt((g ** 2).sum()))
      assert abs(norms[k] - n_o) <= 1e-3 * max(n_o, 1e-6), (k, norms[k], n_o)
      clipped_any |= n_o > 1.0
    p, state = wo.adam_step(p, g_o, state, lr=3e-3, clipnorm=1.0)
    w = m.get_weights()
    for k in p:
      assert rel_err(w[k], p[k]) < 10 * tol, (step, k, rel_err(w[k], p[k]))
  assert clipped_any and opt.iterations == 4
  st = opt.get_state(m)
  for k in p:
    assert rel_err(st['m'][k], state['m'][k]) < 1e-3 and rel_err(st['v'][k], state['v'][k]) < 1e-3, k
  # learning-rate changes (ReduceLROnPlateau, train.py:167-171) take effect on the next step
  opt.learning_rate = 0.0
  w0 = m.get_weights()
  m.train_step(data)
  w1 = m.get_weights()
  assert all(np.array_equal(w0[k], w1[k]) for k in w0)


def test_adam_bf16_tier_trains():
  """bf16 tier: the update runs on the fp32 master weights and the packed bf16 copies follow; the loss goes down."""
  from wavenets_b200 import WaveNet
  from wavenets_b200.optimizers import Adam
  kw = dict(channels=64, blocks=3, layers_per_block=1, dilation_bound=8, skip_channels=64, final_layers_channels=[64])
  x, _ = make_inputs(4, 400, 0, seed=3)
  m = WaveNet(**kw, precision='bf16')
  m.compile(optimizer=Adam(learning_rate=2e-3, clipnorm=1.0))
  m.build(x[:, :-1].shape)
  losses, means = [], []
  for _ in range(12):
    means.append(m.train_step(x)['loss'])
    losses.append(m.last_step_logs['loss'])
  assert losses[-1] < 0.97 * losses[0] and all(b < a for a, b in zip(losses, losses[1:])), losses
  # the returned 'loss' is the running mean of the loss tracker (model.py:166,340-348) until reset_metrics()
  assert all(abs(mu - np.mean(losses[:i + 1])) <= 1e-6 * abs(mu) for i, mu in enumerate(means))
  m.reset_metrics()
  assert m.test_step(x)['loss'] < losses[0]
  assert m.train_step(x)['loss'] != means[-1] and len(m.metrics) == 1


def test_checkpoint_round_trip(tmp_path):
  """Weights-only .npz in Keras layouts; a fresh model restored from it reproduces the loss bit for bit."""
  from wavenets_b200 import WaveNet, checkpoint as ck
  kw = SMALL_MODELS['cond_skip']
  x, cond = make_inputs(2, 64, COND_IN)
  m = WaveNet(**kw)
  m.build((x[:, :-1].shape, cond.shape))
  m.handle.glorot_init(seed=5, bias_std=0.02)
  path = ck.save_weights(m, str(tmp_path / ck.checkpoint_name(3, 5e-4)))
  m2 = WaveNet(**kw)
  m2.build((x[:, :-1].shape, cond.shape))
  ck.load_weights(m2, path)
  assert m.test_step((x, cond))['loss'] == m2.test_step((x, cond))['loss']
  z = np.load(path)
  assert z['block0__dil0__kernel'].shape == (2, 8, 20) and z['mapping0__kernel'].shape == (COND_IN, 4)
  m3 = WaveNet(**{**kw, 'blocks': 2})
  m3.build((x[:, :-1].shape, cond.shape))
  with pytest.raises(ValueError):
    ck.load_weights(m3, path)


def test_train_step_deferred_matches_train_step():
  """Deferred logs (read one call later, pinned D2H + event) carry exactly the values of the synchronous call."""
  from wavenets_b200 import WaveNet
  from wavenets_b200.metrics import MeanSquaredError
  kw = SMALL_MODELS['l2']
  x, cond = make_inputs(2, 80, COND_IN)
  m = WaveNet(**kw)
  m.compile(metrics=[MeanSquaredError()])
  m.build((x[:, :-1].shape, cond.shape))
  ref = m.train_step((x, cond))
  m.reset_metrics()
  m._sample_calls -= 1                       # same Philox stream for the sampled-waveform metric
  handles = [m.train_step_deferred((x, cond)) for _ in range(3)]
  outs = [h.result() for h in handles]
  assert outs[0]['loss'] == ref['loss'] and outs[0]['reg_loss'] == ref['reg_loss']
  assert outs[0]['mean_squared_error'] == ref['mean_squared_error']
  # no optimizer: every step has the same loss, and so has the running mean
  assert outs[1]['loss'] == ref['loss'] and abs(outs[2]['loss'] - ref['loss']) <= 1e-12 * abs(ref['loss'])
  assert m.last_step_logs['loss'] == ref['loss'] and m.loss_tracker.count == 3
  assert handles[0].result() is outs[0]
