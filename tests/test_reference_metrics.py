"""The oracle's metric values against the reference's own metric closures (train.py:102-121):
`tests/golden/metrics.npz` holds what `accurate_MAE / accurate_RMSE / accurate_MSE / nMAE` - cut out of
`/root/reference/train.py` unmodified - return on (y_true, y_pred) pairs, including predictions that cancel a
target exactly, ratings of exactly 0 and batches without any target (`tests/golden/make_metrics_golden.py`).
Keras reports the mean of the [B] vector a metric returns [3P]; `RefModel._values` must give those means."""
import os

import numpy as np
import pytest

from oracle import ref_model
from tests.conftest import GOLDEN

GOLD = np.load(os.path.join(GOLDEN, "metrics.npz"))


@pytest.mark.parametrize("k", range(int(GOLD["n_cases"])))
def test_oracle_metrics_equal_the_reference_closures(k):
    t, y, rr = GOLD["c%d/y_true" % k], GOLD["c%d/y_pred" % k], float(GOLD["c%d/rating_range" % k])
    B, N = t.shape
    m = ref_model.RefModel(1, 4, N, B, use_causal_info=False, dtype=np.float32, rng=np.random.RandomState(0))
    m.compile("adagrad", "mean_squared_error", rating_range=rr)
    loss, mae, acc_mae, nmae, acc_rmse, acc_mse = m._values(y, t)
    want = {n: GOLD["c%d/%s" % (k, n)] for n in ("accurate_MAE", "accurate_RMSE", "accurate_MSE", "nMAE")}
    for got, name in ((acc_mae, "accurate_MAE"), (acc_rmse, "accurate_RMSE"), (acc_mse, "accurate_MSE"), (nmae, "nMAE")):
        ref = np.mean(want[name], dtype=np.float32)
        if np.isnan(ref):                       # no target in the batch: 0 / 0, as in the reference
            assert np.isnan(got), name
        else:
            assert got == pytest.approx(float(ref), rel=2e-6), name
    # 'mae' and the MSE loss are Keras' own per-row means [3P]
    assert mae == pytest.approx(float(np.mean(np.abs(y - t))), rel=2e-6, abs=1e-12)
    assert loss == pytest.approx(float(np.mean(np.square(y - t))), rel=2e-6, abs=1e-12)
