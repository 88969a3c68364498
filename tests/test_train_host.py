"""`train.run` host logic (no GPU): the epoch loop's early stopping (train.py:160-177), which
weights are kept and saved, the evaluate / manual-RMSE test procedures (train.py:202-254). The model
is a scripted stand-in (validation metric per epoch given up front); the reader is the real one on
a golden dataset with the NumPy stream kept on the host."""
import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import train as ocf_train
from tests.helpers import product_reader


class _History(object):
    def __init__(self, d):
        self.history = d


class _FakeNet(object):
    metrics_names = ["loss", "mean_absolute_error", "accurate_MAE", "nMAE", "accurate_RMSE", "accurate_MSE"]

    def __init__(self, script):
        self.script, self.epoch, self.saved, self.weights_set = list(script), 0, [], None
        self.fit_calls, self.logged, self.batches = [], 0, []

    def compile(self, **kw):
        self.compiled = kw

    def fit_generator(self, gen, steps, validation_data=None, validation_steps=None, verbose=0):
        self.fit_calls.append((int(steps), int(validation_steps)))
        v = self.script[self.epoch]
        self.epoch += 1
        return _History({"loss": [1.0], "accurate_MSE": [2.0], "val_accurate_MSE": [v], "val_loss": [v / 10]})

    def get_weights(self):
        return [np.array([self.epoch], dtype=np.float32)]

    def set_weights(self, w):
        self.weights_set = int(w[0][0])

    def save(self, path):
        self.saved.append(path)

    def evaluate_generator(self, gen, steps):
        self.eval_steps = int(steps)
        return [0.1, 0.2, 0.3, 0.4, 0.5, 0.6]

    def test_on_batch(self, batch, sync=True):
        self.batches.append(batch)
        self.logged += 1

    def steps_logged(self):
        return self.logged

    def read_metrics(self, first, count):
        rec = np.zeros((1, 8), dtype=np.float32)
        rec[0, 6] = 4.0 * self.batches[-1].target_count          # pretend every target is off by 2
        return rec


def _reference_early_stopping(vals, patience):
    """train.py:160-177, restated: returns (best_epoch, epochs_run, epochs at which the model is saved)."""
    min_loss, best_epoch, saves = None, 0, []
    for i, v in enumerate(vals):
        if min_loss is None:
            min_loss = v
        elif min_loss > v:
            min_loss, best_epoch = v, i
            saves.append(i)
        elif i - best_epoch > patience:
            return best_epoch, i + 1, saves
    return best_epoch, len(vals), saves


@pytest.mark.parametrize("script,patience", [
    ([5, 4, 4.5, 4.2, 3.9, 4.0, 4.1, 4.3, 9, 9], 1),
    ([5, 6, 7, 8], 0),
    ([5, 6, 4, 6, 6, 6, 3, 9, 9, 9, 9], 2),
    ([3, 2, 1, 0.5], 0),                                         # improves until max_epochs
])
def test_epoch_loop_and_test_procedures(golden_datasets, monkeypatch, tmp_path, script, patience):
    ds = golden_datasets["rev"]
    rd = product_reader(ds, "fixed_split")
    rd.rng_on_device = False
    net = _FakeNet(script)

    class _Owner(object):
        model = net

    monkeypatch.setattr(ocf_train, "omni_model", lambda *a, **k: _Owner())
    cfg = ocf_train.TrainConfig(max_epochs=len(script), batch_size=4, patience=patience,
                                model_save_path=str(tmp_path) + "/", num_hidden_units=8)
    res = ocf_train.run(cfg, reader=rd, rating_range=4.0, verbose=0)
    best, epochs_run, saves = _reference_early_stopping(script, patience)
    assert res["best_epoch"] == best and res["epochs_run"] == epochs_run
    assert res["val_history"] == [float(v) for v in script[:epochs_run]]
    # steps per epoch are floor(n / B) - 1 for train and valid (train.py:157-158)
    assert net.fit_calls[0] == (rd.train_set_size // 4 - 1, rd.val_set_size // 4 - 1)
    # one save per strict improvement after the first epoch (train.py:164-169) + the final best model (:193)
    assert len(net.saved) == len(saves) + 1
    assert [int(p.rsplit("_epoch_", 1)[1].split("_")[0]) for p in net.saved[:-1]] == [i + 1 for i in saves]
    # the tested model is the best-validation one
    assert net.weights_set == best + 1
    # fixed-split test procedures: evaluate_generator over floor(n/B) - 1 steps, manual RMSE over floor(n/B) batches
    assert net.eval_steps == rd.test_set_size // 4 - 1
    assert len(net.batches) == rd.test_set_size // 4
    assert res["test"]["accurate_MSE"] == 0.6
    assert res["manual_test_rmse"] == pytest.approx(2.0)
    rd.close()
