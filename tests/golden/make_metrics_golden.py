#!/usr/bin/env python
"""Golden vectors for the four custom metrics, produced by the reference's own closures (train.py:102-121).

    python tests/golden/make_metrics_golden.py        # needs /root/reference; writes tests/golden/metrics.npz

`train.py` cannot be imported (it is a script over Keras / TensorFlow / h5py and a missing JSON file), so the
four `def`s are cut out of its source with `ast` and executed unmodified in a namespace where the third-party
primitives they call are NumPy restatements of their documented behaviour [3P, Keras 2.0.4 / TF 1.3]:
  tf.count_nonzero(x, dtype=tf.float32) -> float32(np.count_nonzero(x))     tf.sqrt -> np.sqrt
  metrics.mae(y_true, y_pred) -> mean(|y_pred - y_true|, axis=-1)           metrics.mse -> mean((y_pred - y_true)^2, axis=-1)
What the fixture pins is therefore the reference's OWN composition: the count over the whole batch of
`y_true + y_pred != 0`, the `* num_items * batch_size / count` rescaling of the per-row means, the square root
taken per row, the division by `rating_range`."""
import ast
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ("accurate_MAE", "accurate_RMSE", "accurate_MSE", "nMAE")


def reference_metrics(num_items, batch_size, rating_range):
    with open("/root/reference/train.py") as f:
        tree = ast.parse(f.read())
    defs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in NAMES]
    assert sorted(d.name for d in defs) == sorted(NAMES)
    tf = types.SimpleNamespace(float32=np.float32, sqrt=np.sqrt,
                               count_nonzero=lambda x, dtype=None: np.float32(np.count_nonzero(x)))
    metrics = types.SimpleNamespace(mae=lambda t, p: np.mean(np.abs(p - t), axis=-1),
                                    mse=lambda t, p: np.mean(np.square(p - t), axis=-1))
    ns = {"tf": tf, "metrics": metrics, "num_items": num_items, "batch_size": batch_size, "rating_range": rating_range}
    exec(compile(ast.Module(body=defs, type_ignores=[]), "train.py", "exec"), ns)
    return {n: ns[n] for n in NAMES}


def main():
    rs = np.random.RandomState(0)
    store = {}
    k = 0
    for B, N, rr in ((4, 9, 4.0), (16, 50, 4.5), (8, 100, 20.0), (1, 7, 1.0)):
        fns = reference_metrics(N, B, rr)
        for density, mask_value in ((0.3, -1.0), (0.05, -1.0), (0.6, 1.0), (0.0, -1.0)):
            present = rs.random_sample((B, N)) < density
            t = np.where(present, rs.choice([0.0, 0.5, 1, 2, 3.5, 5], size=(B, N)), 0.0).astype(np.float32)
            full = rs.normal(size=(B, N)).astype(np.float32) * 3
            y = (np.where(present, mask_value, 0.0).astype(np.float32) * full).astype(np.float32)
            if density > 0.2:
                i, j = np.argwhere(present)[0]
                y[i, j] = -t[i, j]                       # prediction cancels the target: the entry is NOT counted
                i, j = np.argwhere(present)[1]
                t[i, j] = 0.0                             # a rating of exactly 0 (Jester) is counted through y
            with np.errstate(divide="ignore", invalid="ignore"):
                for name, fn in fns.items():
                    store["c%d/%s" % (k, name)] = np.asarray(fn(t, y), dtype=np.float32)
            store["c%d/y_true" % k], store["c%d/y_pred" % k] = t, y
            store["c%d/rating_range" % k] = np.asarray(rr)
            k += 1
    store["n_cases"] = np.asarray(k)
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **store)
    print(k, "cases")


if __name__ == "__main__":
    main()
