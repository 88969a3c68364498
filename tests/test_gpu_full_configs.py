"""GPU parity of EVERY BASELINE.json config at its FULL size, with the random split drawn on the device.

`tests/test_gpu_xl_sizes.py` covers configs[0] and configs[2] with host-side draws and the product's own
batches densified on the host. Here the oracle side is completely independent of the product:

  * batches: `oracle/ref_batches.batch_stream` (the restatement of `data_reader.py:314-419`, pinned bit-exact
    against the reference's own reader) over a lazily built dict view of the same synthetic split, drawing from
    its own `RandomState(seed)`;
  * arithmetic: `oracle/ref_model.RefModel`, dense float32 NumPy;
  * the product runs `data_reader(..., rng_on_device=True)`: MT19937 stream, cdf and keep flags on the GPU.

Per config: two train steps (the second one sees updated weights and optimizer state), all weights afterwards,
one fixed-split validation batch. Settings follow bench.py's WORKLOADS:
  jester    train_jester.py:19-32,39,61,69-78: users x 100 items, 2 x 256 tanh, causal aux input, masks +1,
            0.0 ratings present, RMSprop (dense rule: every parameter moves every step)
  ml20m     nested-DAE shape, 2 x 512 (what `omni_model` can express) and the width list [1000, 500] (superset:
            the encoder and decoder kernels have different padded widths -> two work-list passes)
  netflix   17 770 item rows x 480 189 user columns, H = 1000 (HP = 1024, 8 float4 per lane in K2/K3/K4b)
"""
import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import optimizers
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
from omnidirectional_collaborative_filtering_b200.model import omni_model
from oracle import ref_batches, ref_model
from tests.helpers import cached_split, mem_available_gb, oracle_data_from_split
from tests.test_gpu_model import _close

pytestmark = pytest.mark.gpu

B = 128


def _weights_close(got, want, lr, dense_rule):
    """Weights after two steps. Elements whose gradient sits at fp32 rounding level may take a normalised step
    (Adagrad / RMSprop divide by the gradient's own magnitude) of a different size on the two sides: allow
    < 0.05 % of such elements, each bounded by two steps of lr; everything else within 1e-3 relative."""
    for g, w in zip(got, want):
        g64, w64 = np.asarray(g, dtype=np.float64), np.asarray(w, dtype=np.float64)
        diff = np.abs(g64 - w64)
        bad = diff > (2e-6 + 1e-3 * np.abs(w64))
        assert bad.mean() < 5e-4, "%d of %d weights off" % (bad.sum(), bad.size)
        assert diff.max() <= (7.0 if dense_rule else 2.5) * lr      # RMSprop's first steps are lr / sqrt(1 - rho) long


def _run(shape, reverse, layers, hidden, act, aux, sparsity, pass_through, opt, lr, pdrop, aux_value, seed=23, steps=2):
    fs = cached_split(shape, reverse)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs, rng_on_device=True)
    data = oracle_data_from_split(fs)
    kw = dict(dense_activation=act, use_causal_info=aux is not None, use_both_masks=aux == "both", dropout_probability=pdrop)
    np.random.seed(5)
    om = omni_model(layers, hidden, fs.n_cols, B, auxilliary_mask_type=aux, **kw)
    ref = ref_model.RefModel(layers, hidden, fs.n_cols, B, dtype=np.float32, rng=np.random.RandomState(0), **kw)
    w0 = om.model.get_weights()
    rs = np.random.RandomState(8)
    for i in range(1, len(w0), 2):                   # non-zero biases
        w0[i] = (rs.normal(size=w0[i].shape) * 0.05).astype(np.float32)
    om.model.set_weights(w0)
    ref.set_weights(w0)
    del w0
    ref.dropout_seed = om.dropout_seed
    o = {"adagrad": optimizers.Adagrad, "rmsprop": optimizers.RMSprop}[opt](lr=lr)
    om.model.compile(o, "mean_squared_error", rating_range=fs.rating_range)
    ref.compile(ref_model.RefOptimizer(opt, lr=lr), "mean_squared_error", rating_range=fs.rating_range)
    np.random.seed(seed)
    gen = rd.data_gen(B, sparsity, "train", True, aux, aux_value, pass_through_input_training=pass_through)
    rgen = ref_batches.batch_stream(data, B, sparsity, "train", True, aux, aux_value,
                                    pass_through_input_training=pass_through, rng=np.random.RandomState(seed), vectorised=True)
    for _ in range(steps):
        batch = next(gen)
        got = om.model.train_on_batch(batch)
        feed, targets = next(rgen)
        # the device drew this batch's split: same number of inputs / targets as the reference's reader
        flags = batch.flags
        mask_out = feed[1] if aux is None else feed[2]
        assert int(np.count_nonzero(mask_out)) == (flags.size if pass_through else int((flags == 0).sum()))
        want = ref.train_on_batch(feed, targets)
        del feed, targets
        _close(got, want)
    _weights_close(om.model.get_weights(), ref.get_weights(), lr, dense_rule=opt != "adagrad")
    # fixed-split validation batch (no RNG besides the order permutation, which both sides draw)
    rd.sync_rng()
    np.random.seed(seed + 1)
    vb = next(rd.data_gen(B, None, "valid", True, aux, aux_value))
    vfeed, vt = next(ref_batches.batch_stream(data, B, None, "valid", True, aux, aux_value, rng=np.random.RandomState(seed + 1)))
    _close(om.model.test_on_batch(vb), ref.test_on_batch(vfeed, vt))
    om.model.close()
    rd.close()


def test_jester_config_full_size():
    _run("jester", False, 2, 256, "tanh", "causal", [0.5, 0.5], False, "rmsprop", 0.001, None, 1)


def test_ml20m_two_equal_layers_full_size():
    _run("ml20m", False, 2, 512, "sigmoid", None, [0.5, 0.5], False, "adagrad", 0.005, 0.2, -1)


def test_ml20m_width_list_full_size():
    _run("ml20m", False, 2, [1000, 500], "sigmoid", None, [0.5, 0.5], False, "adagrad", 0.005, 0.2, -1)


def test_netflix_config_full_size():
    if mem_available_gb() < 60:
        pytest.skip("the dense float32 oracle of the Netflix shape needs ~45 GB of host memory")
    _run("netflix", True, 1, 1000, "sigmoid", None, [1.0, 1.0], True, "adagrad", 0.005, 0.2, -1)
