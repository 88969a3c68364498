#!/usr/bin/env python
"""`train.run` end to end on a GPU against the oracle driven through the same epoch loop; prints every
per-epoch value side by side (the pass/fail version is `tests/test_gpu_train.py`).

    python scripts/train_check.py [autorec|omni]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnidirectional_collaborative_filtering_b200 import synthetic, train as ocf_train
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
from tests.helpers import oracle_train_run
from tests.test_train_loop_host import init_model_for, train_config


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "autorec"
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=8)
    cfg = train_config(name)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    np.random.seed(5)
    got = ocf_train.run(cfg, reader=rd, rating_range=fs.rating_range, save_models=False, verbose=0)
    rd.close()
    want = oracle_train_run(fs, cfg, 5, init_model_for(cfg, fs.n_cols))
    worst = 0.0
    for e, (g, w) in enumerate(zip(got["history"], want["history"])):
        for k in sorted(w):
            worst = max(worst, abs(g[k] - w[k]) / max(abs(w[k]), 1e-9))
            print("epoch %d %-22s product %.7f oracle %.7f" % (e + 1, k, g[k], w[k]))
    for k, v in want["test"].items():
        worst = max(worst, abs(got["test"][k] - v) / max(abs(v), 1e-9))
        print("test    %-22s product %.7f oracle %.7f" % (k, got["test"][k], v))
    worst = max(worst, abs(got["manual_test_rmse"] - want["manual_test_rmse"]) / want["manual_test_rmse"])
    print("manual test RMSE          product %.7f oracle %.7f" % (got["manual_test_rmse"], want["manual_test_rmse"]))
    print("train_check[%s]: worst relative difference %.2e -> %s" % (name, worst, "OK" if worst < 1e-3 else "FAIL"))
    return 0 if worst < 1e-3 else 1


if __name__ == "__main__":
    sys.exit(main())
