"""Checkpoints and donor models on the host (train.py:136-145,164-169,181-199): `m.save` /
`load_model` / `save_weights` / `load_weights` round trips (`.npz` bytes under the reference's file names: h5py is absent here), and the two-stage nested
denoising autoencoder of BASELINE configs[3] through `train.run` - stage 1 saves its best model, stage 2
(`load_weights_from`, deeper, outer layers frozen; or `perform_finetuning`) starts from it. The arithmetic is
the oracle's (`helpers.OracleNet`), so everything here runs without a GPU and must match the oracle loop exactly."""
import glob
import os

import numpy as np
import pytest

from oracle import ref_model
from omnidirectional_collaborative_filtering_b200 import synthetic, train as ocf_train
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
from omnidirectional_collaborative_filtering_b200.model import load_model, omni_model
from tests.helpers import OracleNet, oracle_train_run
from tests.test_train_loop_host import assert_same_run, train_config


def test_save_and_load_model_round_trip(tmp_path):
    np.random.seed(3)
    om = omni_model(2, [12, 6], 30, 8, dense_activation="elu", use_causal_info=True, use_both_masks=True,
                    l2_weight_regulatization=0.01, dropout_probability=0.3, auxilliary_mask_type="both")
    path = str(tmp_path / "model_epoch_3_bestValidScore")
    om.model.save(path)
    np.random.seed(11)
    np.random.random_sample(7)
    state = np.random.get_state()
    back = load_model(path)
    after = np.random.get_state()
    assert state[2] == after[2] and np.array_equal(state[1], after[1])        # loading draws nothing from the caller's stream
    assert back.owner.config() == om.config()
    for a, b in zip(back.get_weights(), om.model.get_weights()):
        assert a.dtype == np.float32 and np.array_equal(a, b)
    # weights only (model.py:102-107)
    om.save_weights(str(tmp_path / "w"))
    other = omni_model(2, [12, 6], 30, 8, dense_activation="elu", use_causal_info=True, use_both_masks=True,
                       auxilliary_mask_type="both")
    assert not np.array_equal(other.model.get_weights()[0], om.model.get_weights()[0])
    other.model.load_weights(str(tmp_path / "w"))
    for a, b in zip(other.model.get_weights(), om.model.get_weights()):
        assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        omni_model(1, 12, 30, 8, use_causal_info=False).model.load_weights(str(tmp_path / "w"))   # wrong architecture


def _oracle_backed(cfg, n_cols):
    real = omni_model                  # the product class itself (ocf_train.omni_model may already be patched)

    def make(*a, **k):
        om = real(*a, **k)
        ref = ref_model.RefModel(cfg.numlayers, cfg.num_hidden_units, n_cols, cfg.batch_size,
                                 dense_activation=cfg.activation_type, use_causal_info=cfg.use_causal_info,
                                 dropout_probability=cfg.dropout_probability, dtype=np.float32,
                                 rng=np.random.RandomState(0))
        ref.set_weights(om.model.get_weights())
        ref.dropout_seed = om.dropout_seed
        om.model = OracleNet(ref, owner=om)
        return om
    return make


@pytest.mark.parametrize("finetune", [False, True], ids=["nested-frozen", "finetune"])
def test_two_stage_training_from_a_saved_donor(monkeypatch, tmp_path, finetune):
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=8)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs, rng_on_device=False)
    save_dir = str(tmp_path) + "/"
    # stage 1: one hidden layer, input corruption; the best-validation model is saved (train.py:164-169,193)
    cfg1 = train_config("autorec", max_epochs=2, train_sparsity=[0.5, 0.5], pass_through_input_training=False,
                        num_hidden_units=24, model_save_path=save_dir)
    monkeypatch.setattr(ocf_train, "omni_model", _oracle_backed(cfg1, fs.n_cols))
    np.random.seed(5)
    first = ocf_train.run(cfg1, reader=rd, rating_range=fs.rating_range, save_models=True, verbose=0)
    saved = sorted(glob.glob(save_dir + "*_bestValidScore"), key=os.path.getmtime)
    assert saved and saved[-1].endswith(first["save_name"] + "_bestValidScore")       # the reference's name, no extension
    donor_name = os.path.basename(saved[-1])
    donor_weights = load_model(save_dir + donor_name).get_weights()
    for a, b in zip(donor_weights, first["model"].model.get_weights()):
        assert np.array_equal(a, b)                       # the tested (best) weights are the saved ones
    # stage 2
    layers2 = 1 if finetune else 3
    cfg2 = train_config("autorec", max_epochs=2, train_sparsity=[0.5, 0.5], pass_through_input_training=False,
                        num_hidden_units=24, numlayers=layers2, model_save_path=save_dir, load_weights_from=donor_name,
                        perform_finetuning=finetune)
    monkeypatch.setattr(ocf_train, "omni_model", _oracle_backed(cfg2, fs.n_cols))
    np.random.seed(6)
    got = ocf_train.run(cfg2, reader=rd, rating_range=fs.rating_range, save_models=False, verbose=0)
    om2 = got["model"]
    final = om2.model.get_weights()
    if finetune:
        assert om2.trainable == [True, True]
        assert all(not np.array_equal(a, b) for a, b in zip(final, donor_weights))     # every layer kept learning
    else:
        assert om2.trainable == [False, True, True, False]                              # model.py:158-170
        for i, j in ((0, 0), (1, 1), (6, 2), (7, 3)):
            assert np.array_equal(final[i], donor_weights[j])                           # frozen outer layers = the donor's
        assert all(final[i].std() > 0 or i % 2 for i in (2, 4))

    def init():                                            # the oracle loop starts from the same donor-initialised model
        om = omni_model(layers2, 24, fs.n_cols, cfg2.batch_size, dense_activation="sigmoid", use_causal_info=False,
                        dropout_probability=cfg2.dropout_probability, auxilliary_mask_type=None)
        if finetune:
            om.manually_load_all_weights(donor_weights)
        else:
            om.load_and_fix_for_denoising_autoencoders(donor_weights)
        return om

    want = oracle_train_run(fs, cfg2, 6, init)
    assert_same_run(got, want, rtol=0)
    for a, b in zip(final, want["weights"]):
        assert np.array_equal(a, b)
    rd.close()
