"""GPU parity, model path (kernels K2-K4 through the C ABI) against the NumPy oracle on the same
batches, weights and dropout masks. Tolerances: per-step metrics and weights after a few steps
within 1e-4 relative of the fp32 oracle (the north-star bar is 1e-3 on per-epoch RMSE)."""
import numpy as np
import pytest

from oracle import ref_batches, ref_model
from omnidirectional_collaborative_filtering_b200 import optimizers
from omnidirectional_collaborative_filtering_b200.model import omni_model
from tests.helpers import oracle_data, product_reader

pytestmark = pytest.mark.gpu

RTOL = 2e-4


def _pair(ds, aux, layers, width, act, l2, pdrop, opt, loss, B=8, seed=7):
    N = ds["n_cols"]
    kw = dict(dense_activation=act, use_causal_info=aux is not None, use_both_masks=aux == "both",
              l2_weight_regulatization=l2, dropout_probability=pdrop)
    np.random.seed(seed)
    om = omni_model(layers, width, N, B, auxilliary_mask_type=aux, **kw)
    ref = ref_model.RefModel(layers, width, N, B, dtype=np.float32, rng=np.random.RandomState(0), **kw)
    w = om.model.get_weights()
    rs = np.random.RandomState(seed + 1)
    for i in range(1, len(w), 2):
        w[i] = (rs.normal(size=w[i].shape) * 0.05).astype(np.float32)      # non-zero biases
    om.model.set_weights(w)
    ref.set_weights(w)
    ref.dropout_seed = om.dropout_seed
    kinds = {"adagrad": optimizers.Adagrad(lr=0.05), "rmsprop": optimizers.RMSprop(lr=0.01),
             "adam": optimizers.Adam(lr=0.01, decay=0.01), "sgd": optimizers.SGD(lr=0.1)}
    o = kinds[opt]
    om.model.compile(optimizer=o, loss=loss, rating_range=4.0)
    ref.compile(ref_model.RefOptimizer(opt, lr=o.lr, epsilon=o.epsilon, decay=o.decay, rho=o.p1 or 0.9,
                                       beta_1=o.p1 or 0.9, beta_2=o.p2 or 0.999), loss, rating_range=4.0)
    return om, ref


def _close(a, b, rtol=RTOL, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def _close_weights(a, b, lr):
    """Weights after a few steps. Adagrad/RMSprop/Adam normalise the step by the gradient's own
    magnitude, so an element whose gradient sits at fp32 rounding-noise level can move by a
    fraction of lr differently on the two sides; allow a handful of such elements (< 0.1 %),
    bounded by the step size, and hold everything else to 1e-3 relative."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    bad = np.abs(a - b) > (2e-6 + 1e-3 * np.abs(b))
    assert bad.mean() < 1e-3, "%d of %d weights off" % (bad.sum(), bad.size)
    assert np.max(np.abs(a - b)) <= 3 * lr


CASES = [
    # aux, layers, width, act, l2, pdrop, opt, loss, pass_through
    (None, 1, 12, "sigmoid", None, None, "adagrad", "mean_squared_error", True),
    ("dropout", 1, 200, "tanh", None, 0.2, "adagrad", "mean_squared_error", False),
    ("causal", 2, 40, "elu", 0.01, None, "rmsprop", "mean_squared_error", False),
    ("both", 3, 130, "selu", None, 0.5, "adam", "mean_absolute_error", True),
    ("zeros", 2, 24, "softplus", 0.001, None, "sgd", "mean_squared_error", False),
    (None, 1, 300, "relu", None, None, "adam", "mean_squared_error", True),
    ("dropout", 2, [36, 20], "linear", None, 0.3, "adagrad", "mean_squared_error", False),
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_train_steps_match_oracle(golden_datasets, case):
    aux, layers, width, act, l2, pdrop, opt, loss, pt = case
    ds = golden_datasets["rev"]
    om, ref = _pair(ds, aux, layers, width, act, l2, pdrop, opt, loss)
    rd = product_reader(ds, "fixed_split")
    data = oracle_data(ds, "fixed_split")
    np.random.seed(3)
    gen = rd.data_gen(8, [0.2, 0.9], "train", True, aux, -1, pass_through_input_training=pt)
    # the oracle draws from its own RandomState(3): same numbers as the global stream seeded 3,
    # without the two lazy generators interleaving their draws
    rgen = ref_batches.batch_stream(data, 8, [0.2, 0.9], "train", True, aux, -1, pass_through_input_training=pt,
                                    rng=np.random.RandomState(3))
    for step in range(5):
        got = om.model.train_on_batch(next(gen))
        feed, targets = next(rgen)
        want = ref.train_on_batch(feed, targets)
        _close(got, want)
    for g, w in zip(om.model.get_weights(), ref.get_weights()):
        _close_weights(g, w, om.model.optimizer.lr)
    # evaluation on a fixed-split batch + predict + full-catalogue scores
    np.random.seed(4)
    vgen = rd.data_gen(8, None, "valid", True, aux, -1)
    rvgen = ref_batches.batch_stream(data, 8, None, "valid", True, aux, -1, rng=np.random.RandomState(4))
    vb = next(vgen)
    vfeed, vt = next(rvgen)
    _close(om.model.test_on_batch(vb), ref.test_on_batch(vfeed, vt))
    _close(om.model.predict(vb), ref.predict(vfeed), rtol=1e-3, atol=1e-5)
    # tensor-core scoring rounds its operands to tf32 (2^-11): see tests/test_gpu_score.py
    want_scores = ref.score(vfeed)
    _close(om.model.score(vb), want_scores, rtol=2e-3, atol=2e-3 * max(1.0, float(np.abs(want_scores).max())))
    rd.close()


@pytest.mark.parametrize("case", [CASES[1], CASES[2], CASES[3], CASES[6]], ids=["adagrad", "rmsprop-l2", "adam-mae", "widths"])
def test_row_parallel_gradient_path_matches_oracle(golden_datasets, case):
    """The data-parallel step of one rank (world 1: backward into the gradient arena, streaming
    optimizer pass, no all-reduce) is the same math as the fused step."""
    from omnidirectional_collaborative_filtering_b200 import _lib
    aux, layers, width, act, l2, pdrop, opt, loss, pt = case
    ds = golden_datasets["rev"]
    om, ref = _pair(ds, aux, layers, width, act, l2, pdrop, opt, loss)
    om.model.native = (None, _lib.PAR_ROWS)
    rd = product_reader(ds, "fixed_split")
    data = oracle_data(ds, "fixed_split")
    np.random.seed(3)
    gen = rd.data_gen(8, [0.2, 0.9], "train", True, aux, -1, pass_through_input_training=pt)
    rgen = ref_batches.batch_stream(data, 8, [0.2, 0.9], "train", True, aux, -1, pass_through_input_training=pt,
                                    rng=np.random.RandomState(3))
    for step in range(5):
        got = om.model.train_on_batch(next(gen))
        feed, targets = next(rgen)
        _close(got, ref.train_on_batch(feed, targets))
    for g, w in zip(om.model.get_weights(), ref.get_weights()):
        _close_weights(g, w, om.model.optimizer.lr)
    vb = next(rd.data_gen(8, None, "valid", False, aux, -1))
    vfeed, vt = next(ref_batches.batch_stream(data, 8, None, "valid", False, aux, -1, rng=np.random.RandomState(4)))
    _close(om.model.test_on_batch(vb), ref.test_on_batch(vfeed, vt))
    rd.close()


SYN_CASES = [
    # shape, reverse, B, aux, layers, width, opt, l2, pdrop, pass_through   (no row repeats a column here:
    # the weight update takes its work list from the batch-side counting sort, not from the CSC scan)
    ("small", True, 64, None, 1, 48, "adagrad", None, 0.2, True),
    ("small", False, 160, "dropout", 1, 130, "adagrad", None, None, False),
    ("small", True, 96, "both", 2, 40, "rmsprop", 0.01, None, False),
    ("tiny", True, 8, "causal", 1, 12, "adam", None, 0.3, True),
]


@pytest.mark.parametrize("case", SYN_CASES, ids=[str(i) for i in range(len(SYN_CASES))])
def test_train_steps_match_oracle_synthetic(case):
    from omnidirectional_collaborative_filtering_b200 import synthetic
    from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
    shape, rev, B, aux, layers, width, opt, l2, pdrop, pt = case
    fs = synthetic.make_fixed_split(shape, reverse_user_item_data=rev, seed=21)
    dicts = synthetic.to_reference_dicts(fs, raw_col_id=lambda c: c)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    assert rd.store("train").info()["has_dups"] == 0
    data = ref_batches.RefData(fs.n_cols, fs.train.n_rows, dicts["unique_cols"], eval_mode="fixed_split",
                               train=dicts["train"], valid=tuple(dicts["valid"]), test=tuple(dicts["test"]))
    om, ref = _pair({"n_cols": fs.n_cols}, aux, layers, width, "sigmoid", l2, pdrop, opt, "mean_squared_error", B=B)
    np.random.seed(13)
    gen = rd.data_gen(B, [0.3, 0.9], "train", True, aux, -1, pass_through_input_training=pt)
    rgen = ref_batches.batch_stream(data, B, [0.3, 0.9], "train", True, aux, -1, pass_through_input_training=pt,
                                    rng=np.random.RandomState(13), vectorised=True)
    for step in range(4):
        got = om.model.train_on_batch(next(gen))
        feed, targets = next(rgen)
        _close(got, ref.train_on_batch(feed, targets))
    for g, w in zip(om.model.get_weights(), ref.get_weights()):
        _close_weights(g, w, om.model.optimizer.lr)
    rd.close()


def test_frozen_layers_and_transfer(golden_datasets):
    ds = golden_datasets["fwd"]
    N = ds["n_cols"]
    np.random.seed(1)
    donor = omni_model(1, 16, N, 8, dense_activation="sigmoid", use_causal_info=False)
    new = omni_model(3, 16, N, 8, dense_activation="sigmoid", use_causal_info=False)
    new.load_and_fix_for_denoising_autoencoders(donor)
    assert new.trainable == [False, True, True, False]
    before = new.model.get_weights()
    assert np.array_equal(before[0], donor.model.get_weights()[0])
    assert np.array_equal(before[-2], donor.model.get_weights()[-2])
    rd = product_reader(ds, "fixed_split")
    np.random.seed(2)
    gen = rd.data_gen(8, [0.5, 0.5], "train", True, None, -1)
    new.model.compile(optimizers.Adagrad(lr=0.05))
    for _ in range(3):
        new.model.train_on_batch(next(gen))
    after = new.model.get_weights()
    for i in (0, 1, 6, 7):
        assert np.array_equal(before[i], after[i])
    for i in (2, 3, 4, 5):
        assert not np.array_equal(before[i], after[i])
    rd.close()


def test_epoch_metrics_match_oracle(golden_datasets):
    """fit_generator / evaluate_generator over whole epochs (train.py:150-158 semantics)."""
    ds = golden_datasets["fwd"]
    om, ref = _pair(ds, None, 1, 64, "sigmoid", None, 0.2, "adagrad", "mean_squared_error", B=4)
    rd = product_reader(ds, "fixed_split")
    data = oracle_data(ds, "fixed_split")
    for epoch in range(3):
        np.random.seed(100 + epoch)
        tg = rd.data_gen(4, [1.0, 1.0], "train", True, None, -1, pass_through_input_training=True)
        vg = rd.data_gen(4, [1.0, 1.0], "valid", True, None, -1)
        hist = om.model.fit_generator(tg, np.floor(rd.train_set_size / 4) - 1, validation_data=vg,
                                      validation_steps=np.floor(rd.val_set_size / 4) - 1, verbose=0).history
        rrng = np.random.RandomState(100 + epoch)      # train stream first, then valid, like the product
        rtg = ref_batches.batch_stream(data, 4, [1.0, 1.0], "train", True, None, -1, pass_through_input_training=True, rng=rrng)
        rvg = ref_batches.batch_stream(data, 4, [1.0, 1.0], "valid", True, None, -1, rng=rrng)
        rhist = ref.fit_generator(rtg, np.floor(data.train_set_size / 4) - 1, validation_data=rvg,
                                  validation_steps=np.floor(data.val_set_size / 4) - 1)
        for k, v in rhist.items():
            _close(hist[k][-1], v[-1], rtol=1e-3)
    rd.close()
