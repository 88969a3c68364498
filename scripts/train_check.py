#!/usr/bin/env python
"""`train.run` end to end on a GPU (not part of the test-suite yet: written when the round's GPU
budget was spent - run it first thing next round and move it into tests/ once green):
three epochs of the reference's training script on a small synthetic fixed-split dataset, against
the oracle driven through the same epoch loop (per-epoch train/valid metrics within 1e-3).

    python scripts/train_check.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_batches, ref_model
from omnidirectional_collaborative_filtering_b200 import synthetic, train as ocf_train
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader


def main():
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=8)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    B, H, epochs = 32, 48, 3
    cfg = ocf_train.TrainConfig(max_epochs=epochs, batch_size=B, patience=5, num_hidden_units=H,
                                dropout_probability=0.2, model_save_path="/tmp/ocf_train_check/")
    np.random.seed(5)
    res = ocf_train.run(cfg, reader=rd, rating_range=fs.rating_range, save_models=False, verbose=0)
    om = res["model"]

    # the oracle through the same loop: same initial weights are not recoverable after training, so
    # re-create the model from the same seed and copy its weights before any step
    dicts = synthetic.to_reference_dicts(fs, raw_col_id=lambda c: c)
    data = ref_batches.RefData(fs.n_cols, fs.train.n_rows, dicts["unique_cols"], eval_mode="fixed_split",
                               train=dicts["train"], valid=tuple(dicts["valid"]), test=tuple(dicts["test"]))
    np.random.seed(5)
    from omnidirectional_collaborative_filtering_b200.model import omni_model
    twin = omni_model(1, H, fs.n_cols, B, dense_activation="sigmoid", use_causal_info=False, dropout_probability=0.2)
    ref = ref_model.RefModel(1, H, fs.n_cols, B, dense_activation="sigmoid", use_causal_info=False,
                             dropout_probability=0.2, dtype=np.float32)
    ref.set_weights(twin.model.get_weights())
    ref.dropout_seed = twin.dropout_seed
    ref.compile(ref_model.RefOptimizer("adagrad", lr=0.005), "mean_squared_error", rating_range=fs.rating_range)
    rng = np.random.RandomState()
    rng.set_state(np.random.get_state())          # the stream right after the model's initialisation
    worst = 0.0
    for e in range(epochs):
        tg = ref_batches.batch_stream(data, B, [1.0, 1.0], "train", True, None, -1, pass_through_input_training=True,
                                      rng=rng, vectorised=True)
        vg = ref_batches.batch_stream(data, B, [1.0, 1.0], "valid", True, None, -1, rng=rng, vectorised=True)
        h = ref.fit_generator(tg, np.floor(data.train_set_size / B) - 1, validation_data=vg,
                              validation_steps=np.floor(data.val_set_size / B) - 1)
        for k in ("accurate_MSE", "val_accurate_MSE", "accurate_RMSE", "val_accurate_RMSE", "loss"):
            got, want = res["history"][e][k], h[k][-1]
            worst = max(worst, abs(got - want) / max(abs(want), 1e-9))
            print("epoch %d %-18s product %.6f oracle %.6f" % (e + 1, k, got, want))
    print("train_check: worst relative difference %.2e -> %s" % (worst, "OK" if worst < 1e-3 else "FAIL"))
    print("test:", res["test"], "manual RMSE", res["manual_test_rmse"])
    rd.close()
    return 0 if worst < 1e-3 else 1


if __name__ == "__main__":
    sys.exit(main())
