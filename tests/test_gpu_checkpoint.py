"""Checkpoints on the GPU (train.py:136-145,164-169,181-199): `train.run(save_models=True)` writes
the best-validation model from device memory, `load_model` puts it back on the device with the same bits, testing
the reloaded model reproduces the run's test metrics, and a second stage (`load_weights_from`: nested denoising AE
with frozen outer layers, or `perform_finetuning`) continues on the device from the saved donor - each against the
oracle loop started from the same donor weights. The host-only twin of this file is `tests/test_checkpoint_host.py`."""
import glob
import os

import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import synthetic, train as ocf_train
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
from omnidirectional_collaborative_filtering_b200.model import load_model, omni_model
from tests.helpers import oracle_train_run
from tests.test_train_loop_host import assert_same_run, train_config

pytestmark = pytest.mark.gpu


def _newest(save_dir, pattern):
    found = sorted(glob.glob(glob.escape(save_dir) + pattern), key=os.path.getmtime)
    assert found, "no checkpoint written under %s" % save_dir
    return found[-1]


@pytest.mark.parametrize("finetune", [False, True], ids=["nested-frozen", "finetune"])
def test_best_model_saved_reloaded_and_continued_on_the_device(tmp_path, finetune):
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=8)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    save_dir = str(tmp_path) + "/"
    # ---- stage 1: one hidden layer with input corruption; the best-validation model is saved ----------------
    cfg1 = train_config("autorec", max_epochs=3, train_sparsity=[0.5, 0.5], pass_through_input_training=False,
                        num_hidden_units=24, model_save_path=save_dir)
    np.random.seed(5)
    first = ocf_train.run(cfg1, reader=rd, rating_range=fs.rating_range, save_models=True, verbose=0)
    path = _newest(save_dir, glob.escape(first["save_name"]) + "_bestValidScore*")
    live = first["model"].model                           # holds the best weights (train.py:191: the tested model)
    # ---- reload: same bits on the device, same test metrics ---------------------------------------------------
    back = load_model(path)
    for a, b in zip(back.get_weights(), live.get_weights()):
        assert a.dtype == np.float32 and np.array_equal(a, b)
    back.compile(ocf_train.Adagrad(lr=cfg1.learning_rate, epsilon=1e-08, decay=0.0), cfg1.model_loss, rating_range=fs.rating_range)
    steps = np.floor(rd.test_set_size / cfg1.batch_size) - 1
    np.random.seed(77)
    again = back.evaluate_generator(rd.data_gen(cfg1.batch_size, None, "test", True, None, -1), steps)
    np.random.seed(77)
    ref_vals = live.evaluate_generator(rd.data_gen(cfg1.batch_size, None, "test", True, None, -1), steps)
    assert again == ref_vals                              # bit-identical: same weights, fixed-order reductions
    # a reloaded model carries no optimizer state (train.py:183-189 strips it): its first step equals the first
    # step of a fresh model given the same weights
    fresh = omni_model(1, 24, fs.n_cols, cfg1.batch_size, dense_activation="sigmoid", use_causal_info=False,
                       dropout_probability=cfg1.dropout_probability, auxilliary_mask_type=None)
    fresh.model.set_weights(live.get_weights())
    fresh.dropout_seed = back.owner.dropout_seed
    fresh.model.compile(ocf_train.Adagrad(lr=cfg1.learning_rate, epsilon=1e-08, decay=0.0), cfg1.model_loss, rating_range=fs.rating_range)
    rd.sync_rng()
    np.random.seed(78)
    b1 = next(rd.data_gen(cfg1.batch_size, [0.5, 0.5], "train", True, None, -1))
    got1 = back.train_on_batch(b1)
    b1.flags
    want1 = fresh.model.train_on_batch(b1)
    assert got1 == want1
    for a, b in zip(back.get_weights(), fresh.model.get_weights()):
        assert np.array_equal(a, b)
    back.close(); fresh.model.close()
    # ---- stage 2 continues on the device from the saved donor -------------------------------------------------
    donor_name = os.path.basename(path)
    donor_weights = load_model(path).get_weights()
    layers2 = 1 if finetune else 3
    cfg2 = train_config("autorec", max_epochs=2, train_sparsity=[0.5, 0.5], pass_through_input_training=False,
                        num_hidden_units=24, numlayers=layers2, model_save_path=save_dir, load_weights_from=donor_name,
                        perform_finetuning=finetune)
    rd.sync_rng()
    np.random.seed(6)
    got = ocf_train.run(cfg2, reader=rd, rating_range=fs.rating_range, save_models=True, verbose=0)
    final = got["model"].model.get_weights()
    if finetune:
        assert got["model"].trainable == [True, True]
    else:
        assert got["model"].trainable == [False, True, True, False]                     # model.py:158-170
        for i, j in ((0, 0), (1, 1), (6, 2), (7, 3)):
            assert np.array_equal(final[i], donor_weights[j])                           # frozen layers never moved

    def init():
        om = omni_model(layers2, 24, fs.n_cols, cfg2.batch_size, dense_activation="sigmoid", use_causal_info=False,
                        dropout_probability=cfg2.dropout_probability, auxilliary_mask_type=None)
        if finetune:
            om.manually_load_all_weights(donor_weights)
        else:
            om.load_and_fix_for_denoising_autoencoders(donor_weights)
        return om

    want = oracle_train_run(fs, cfg2, 6, init)
    assert_same_run(got, want, rtol=1e-3)
    # the second stage's own best model is on disk too and reloads to the tested weights
    path2 = _newest(save_dir, glob.escape(got["save_name"]) + "_bestValidScore*")
    for a, b in zip(load_model(path2).get_weights(), final):
        assert np.array_equal(a, b)
    rd.close()
