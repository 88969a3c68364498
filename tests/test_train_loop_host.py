"""`train.run` end to end without a GPU: the product's reader, generators, NumPy-stream order, epoch
loop, best-weights bookkeeping and both fixed-split test procedures, with the arithmetic supplied by
the oracle (`helpers.OracleNet` consumes the product's `Batch` objects). Against the oracle driven
through its own restatement of the loop (`helpers.oracle_train_run`: `ref_batches.batch_stream` on one
RandomState) every number must be IDENTICAL - any difference is host logic (which rows, which draws,
which step counts), not rounding. `tests/test_gpu_train.py` is the same comparison with the CUDA
model in place of `OracleNet`."""
import numpy as np
import pytest

from oracle import ref_model
from omnidirectional_collaborative_filtering_b200 import synthetic, train as ocf_train
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
from tests.helpers import OracleNet, oracle_train_run

CONFIGS = {
    # train.py's defaults: plain AutoRec (every rating input and target), no aux input, dropout 0.2
    "autorec": dict(train_sparsity=[1.0, 1.0], pass_through_input_training=True, use_causal_info=False,
                    auxilliary_mask_type=None, dropout_probability=0.2),
    # omnidirectional: reciprocal random split drawn per row, aux mask input
    "omni": dict(train_sparsity=[0.2, 0.9], pass_through_input_training=False, use_causal_info=True,
                 auxilliary_mask_type="dropout", dropout_probability=None, activation_type="tanh"),
}


def train_config(name, **kw):
    base = dict(max_epochs=3, batch_size=32, patience=5, num_hidden_units=48, model_save_path="/tmp/ocf_unused/")
    base.update(CONFIGS[name])
    base.update(kw)
    return ocf_train.TrainConfig(**base)


def init_model_for(cfg, n_cols):
    from omnidirectional_collaborative_filtering_b200.model import omni_model
    return lambda: omni_model(cfg.numlayers, cfg.num_hidden_units, n_cols, cfg.batch_size,
                              dense_activation=cfg.activation_type, use_causal_info=cfg.use_causal_info,
                              use_both_masks=cfg.auxilliary_mask_type == "both",
                              l2_weight_regulatization=cfg.l2_weight_regulatization,
                              dropout_probability=cfg.dropout_probability,
                              auxilliary_mask_type=cfg.auxilliary_mask_type)


def assert_same_run(got, want, rtol):
    assert got["epochs_run"] == len(want["history"]) and got["best_epoch"] == want["best_epoch"]
    for e, (g, w) in enumerate(zip(got["history"], want["history"])):
        for k, v in w.items():
            assert g[k] == pytest.approx(v, rel=rtol, abs=0), "epoch %d %s" % (e + 1, k)
    for k, v in want["test"].items():
        assert got["test"][k] == pytest.approx(v, rel=rtol, abs=0), "test %s" % k
    # the step record carries the batch's squared error as float32
    assert got["manual_test_rmse"] == pytest.approx(want["manual_test_rmse"], rel=max(rtol, 1e-6), abs=0)


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_train_run_host_logic_is_the_oracle_loop(monkeypatch, name):
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=8)
    cfg = train_config(name)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs, rng_on_device=False)
    real_omni = ocf_train.omni_model

    def oracle_backed(*a, **k):
        om = real_omni(*a, **k)                 # same initial weights / dropout seed / stream position
        ref = ref_model.RefModel(cfg.numlayers, cfg.num_hidden_units, fs.n_cols, cfg.batch_size,
                                 dense_activation=cfg.activation_type, use_causal_info=cfg.use_causal_info,
                                 dropout_probability=cfg.dropout_probability, dtype=np.float32,
                                 rng=np.random.RandomState(0))
        ref.set_weights(om.model.get_weights())
        ref.dropout_seed = om.dropout_seed
        om.model = OracleNet(ref)
        return om

    monkeypatch.setattr(ocf_train, "omni_model", oracle_backed)
    np.random.seed(5)
    got = ocf_train.run(cfg, reader=rd, rating_range=fs.rating_range, save_models=False, verbose=0)
    monkeypatch.setattr(ocf_train, "omni_model", real_omni)
    want = oracle_train_run(fs, cfg, 5, init_model_for(cfg, fs.n_cols))
    assert_same_run(got, want, rtol=0)
    rd.close()


def test_train_run_ablation_mode_is_the_oracle_loop(monkeypatch, golden_datasets):
    """eval_mode='ablation' (train.py:85-86,202-213): the row split is drawn before the model is built, validation
    and test batches go through the random split too, one evaluation per test sparsity (scalars accepted)."""
    from tests.helpers import oracle_train_run_ablation, product_reader
    ds = golden_datasets["rev"]
    rd = product_reader(ds, "ablation")
    rd.rng_on_device = False
    cfg = train_config("omni", eval_mode="ablation", batch_size=4, num_hidden_units=10, test_sparsities=[0.0, 0.5, 0.9],
                       val_split=[0.6, 0.2, 0.2])
    N = ds["n_cols"]
    real_omni = ocf_train.omni_model

    def oracle_backed(*a, **k):
        om = real_omni(*a, **k)
        ref = ref_model.RefModel(cfg.numlayers, cfg.num_hidden_units, N, cfg.batch_size, dense_activation=cfg.activation_type,
                                 use_causal_info=cfg.use_causal_info, dropout_probability=cfg.dropout_probability,
                                 dtype=np.float32, rng=np.random.RandomState(0))
        ref.set_weights(om.model.get_weights())
        ref.dropout_seed = om.dropout_seed
        om.model = OracleNet(ref, owner=om)
        return om

    monkeypatch.setattr(ocf_train, "omni_model", oracle_backed)
    np.random.seed(9)
    got = ocf_train.run(cfg, reader=rd, rating_range=4.0, save_models=False, verbose=0)
    monkeypatch.setattr(ocf_train, "omni_model", real_omni)
    want = oracle_train_run_ablation(ds["ablation"], ds["unique_cols"], ds["unique_rows"], cfg, 9,
                                     init_model_for(cfg, N), 4.0)
    assert got["best_epoch"] == want["best_epoch"] and len(got["history"]) == len(want["history"])
    for g, w in zip(got["history"], want["history"]):
        assert g == w
    assert set(got["test"]) == set(want["test"])
    for s, vals in want["test"].items():
        assert got["test"][s] == vals
    rd.close()


def test_csv_to_metrics_pipeline_is_the_oracle_loop(monkeypatch, tmp_path):
    """The whole drop-in story without a GPU: ratings CSV -> `splitter.split_data` -> the reference's files ->
    `data_reader(use_json=True)` (native ingest) -> `train.run`; against the oracle loop over the same files read
    with json.load. `tests/test_gpu_train.py` has the CUDA twin."""
    from tests.helpers import files_pipeline
    d, n_items, n_rows, data = files_pipeline(tmp_path)
    cfg = train_config("autorec", batch_size=8, num_hidden_units=16, max_epochs=2, reverse_user_item_data=False)
    rd = data_reader(n_items, n_rows, d, use_json=True, eval_mode="fixed_split", rng_on_device=False)
    real_omni = ocf_train.omni_model

    def oracle_backed(*a, **k):
        om = real_omni(*a, **k)
        ref = ref_model.RefModel(cfg.numlayers, cfg.num_hidden_units, n_items, cfg.batch_size, dense_activation=cfg.activation_type,
                                 use_causal_info=cfg.use_causal_info, dropout_probability=cfg.dropout_probability,
                                 dtype=np.float32, rng=np.random.RandomState(0))
        ref.set_weights(om.model.get_weights())
        ref.dropout_seed = om.dropout_seed
        om.model = OracleNet(ref, owner=om)
        return om

    monkeypatch.setattr(ocf_train, "omni_model", oracle_backed)
    np.random.seed(5)
    got = ocf_train.run(cfg, reader=rd, rating_range=4.5, save_models=False, verbose=0)
    monkeypatch.setattr(ocf_train, "omni_model", real_omni)
    want = oracle_train_run(None, cfg, 5, init_model_for(cfg, n_items), data=data, n_cols=n_items, rating_range=4.5)
    assert_same_run(got, want, rtol=0)
    rd.close()
