"""Running-mean metrics with the Keras surface the reference uses (train.py:227 `tf.keras.metrics.MeanSquaredError()`,
model.py:166-168 `tf.keras.metrics.Mean`).  Values are produced on the GPU inside train_step / test_step."""


class Mean:
  def __init__(self, name='mean'):
    self.name = name
    self.reset_state()

  def reset_state(self):
    self.total, self.count = 0.0, 0

  def update_state(self, value, sample_weight=None):
    self.total += float(value)
    self.count += 1

  def result(self):
    return self.total / self.count if self.count else 0.0


class MeanSquaredError(Mean):
  """Mean over steps of mean((y_true - y_pred)^2); inside the model the per-step value comes from
  `wn_sample_last_step` (y_pred = a waveform sampled from the predictive distribution, model.py:338-346)."""

  def __init__(self, name='mean_squared_error'):
    super().__init__(name=name)
