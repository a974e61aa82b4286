"""End-to-end: the reference's train.py flow on synthetic audio with every stage on the GPU (examples/train_synthetic.py):
input pipeline -> model -> Adam + clipnorm -> metrics -> checkpoint -> resume -> generation."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_synthetic_runs_learns_and_resumes(tmp_path, monkeypatch, capsys):
  sys.path.insert(0, os.path.join(ROOT, 'examples'))
  import train_synthetic
  res = str(tmp_path / 'run')
  monkeypatch.setattr(sys, 'argv', ['train_synthetic.py', '--steps', '60', '--results', res])
  logs1 = train_synthetic.main()
  out1 = capsys.readouterr().out
  assert 'saved' in out1 and 'generated (2, 1000, 1)' in out1
  assert logs1['mean_squared_error'] > 0
  monkeypatch.setattr(sys, 'argv', ['train_synthetic.py', '--steps', '20', '--results', res])
  logs2 = train_synthetic.main()
  out2 = capsys.readouterr().out
  assert 'resuming from' in out2 and 'epoch 1' in out2
  first_loss_line = [l for l in out1.splitlines() if l.startswith('step')][0]
  first_loss = float(first_loss_line.split('loss')[1].split()[0])
  assert logs1['loss'] < 0.8 * first_loss          # it learns
  assert logs2['loss'] < 0.9 * first_loss          # and continues from the checkpoint, not from scratch
  assert sorted(os.listdir(res)) == ['weights-e0001-lr0.001.weights.npz', 'weights-e0002-lr0.001.weights.npz']
