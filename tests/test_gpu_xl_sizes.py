"""GPU parity at BASELINE.json's FULL sizes (not the reduced shapes of the other GPU tests).

  * configs[0], ML-1M shape (3 706 item rows x 6 040 user columns, H = 500, B = 128, plain AutoRec): the dense
    NumPy oracle still finishes a step in milliseconds, so train steps, weights and a validation batch are
    compared with it directly.
  * configs[2] as `bench.py` runs it by default (ML-10M shape, rows = items, 71 567 columns, k = 2 input blocks,
    H = 512, reciprocal split [0.5, 0.5], dropout 0.2): two train steps + one validation batch against the oracle
    (a dense [128, 143 134] x [143 134, 512] step is ~1 s on the host), plus the size-independent properties:
    the run is bit-reproducible, and evaluation metrics do not depend on the order of the rows in a batch.
The oracle is fed the product's own batches densified on the host (`helpers.host_densify`); that the batches
themselves are the reference's is `tests/test_gpu_batches.py` / `test_gpu_rng.py`."""
import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import optimizers, synthetic
from omnidirectional_collaborative_filtering_b200.data_reader import Batch, data_reader
from omnidirectional_collaborative_filtering_b200.model import omni_model
from oracle import ref_model
from tests.helpers import host_densify
from tests.test_gpu_model import _close, _close_weights

pytestmark = pytest.mark.gpu


def _models(n_cols, H, B, aux, pdrop, seed, copies=1, opt="adagrad", l2=None, layers=1):
    kw = dict(dense_activation="sigmoid", use_causal_info=aux is not None, dropout_probability=pdrop,
              l2_weight_regulatization=l2)
    out = []
    for _ in range(copies):
        np.random.seed(seed)
        om = omni_model(layers, H, n_cols, B, auxilliary_mask_type=aux, **kw)
        o = optimizers.Adagrad(lr=0.005, epsilon=1e-08, decay=0.0) if opt == "adagrad" else optimizers.Adam(lr=0.001)
        om.model.compile(o, "mean_squared_error", rating_range=4.5)
        out.append(om)
    ref = ref_model.RefModel(layers, H, n_cols, B, dtype=np.float32, rng=np.random.RandomState(0), **kw)
    ref.set_weights(out[0].model.get_weights())
    ref.dropout_seed = out[0].dropout_seed
    ro = (ref_model.RefOptimizer("adagrad", lr=0.005) if opt == "adagrad" else
          ref_model.RefOptimizer("adam", lr=o.lr, epsilon=o.epsilon, decay=o.decay, beta_1=o.p1, beta_2=o.p2))
    ref.compile(ro, "mean_squared_error", rating_range=4.5)
    return out, ref


def _run_against_oracle(fs, H, B, aux, sparsity, pass_through, pdrop, steps, copies=1, opt="adagrad", l2=None, layers=1):
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs, rng_on_device=False)
    oms, ref = _models(fs.n_cols, H, B, aux, pdrop, seed=3, copies=copies, opt=opt, l2=l2, layers=layers)
    np.random.seed(17)
    gen = rd.data_gen(B, sparsity, "train", True, aux, -1, pass_through_input_training=pass_through)
    logs = [[] for _ in oms]
    for _ in range(steps):
        batch = next(gen)
        feed, targets = host_densify(batch)
        want = ref.train_on_batch(feed, targets)
        for om, log in zip(oms, logs):
            got = om.model.train_on_batch(batch)
            log.append(got)
            _close(got, want)
    weights = [om.model.get_weights() for om in oms]
    for g, w in zip(weights[0], ref.get_weights()):
        _close_weights(g, w, 0.005 if opt == "adagrad" else 0.001)
    vb = next(rd.data_gen(B, None, "valid", True, aux, -1))
    vfeed, vt = host_densify(vb)
    got_eval = oms[0].model.test_on_batch(vb)
    _close(got_eval, ref.test_on_batch(vfeed, vt))
    return rd, oms, logs, weights, vb, got_eval


def test_ml1m_config_matches_oracle_at_full_size():
    fs = synthetic.make_fixed_split("ml1m", reverse_user_item_data=True, seed=0)
    assert (fs.train.n_rows, fs.n_cols) == (3706, 6040)
    rd, oms, logs, _, _, _ = _run_against_oracle(fs, H=500, B=128, aux=None, sparsity=[1.0, 1.0], pass_through=True,
                                                  pdrop=0.2, steps=4)
    assert logs[0][-1][0] < logs[0][0][0] * 1.5          # sane losses (the four batches differ; no blow-up)
    rd.close()


def test_ml10m_bench_workload_matches_oracle_and_is_reproducible():
    fs = synthetic.make_fixed_split("ml10m", reverse_user_item_data=True, seed=0)
    assert fs.n_cols == 71567 and fs.train.nnz > 7_900_000
    rd, oms, logs, weights, vb, got_eval = _run_against_oracle(fs, H=512, B=128, aux="dropout", sparsity=[0.5, 0.5],
                                                                pass_through=False, pdrop=0.2, steps=2, copies=2)
    # bit-reproducible: every reduction runs in a fixed order, no float atomics
    assert logs[0] == logs[1]
    for a, b in zip(weights[0], weights[1]):
        assert np.array_equal(a, b)
    # evaluation does not depend on the order of the rows inside a batch (row statistics are per row; only the
    # final sums over rows are re-associated)
    perm = np.random.RandomState(1).permutation(vb.n_rows)
    shuffled = Batch(rd, "fixed", vb.source, vb.rows[perm], None, False, vb.aux_type, vb.aux_value, vb.target_count,
                     False, vb.n_ratings)
    _close(oms[0].model.test_on_batch(shuffled), got_eval, rtol=1e-5)
    rd.close()


def test_ml10m_shape_dense_rule_on_the_lean_update():
    """Adam + L2 at the ML-10M shape: every parameter of the 71 567-column catalogue moves every step, so the row
    update enumerates all (column, array) pairs, and at this size it is the lean variant (task cursor, evict-first
    state and stores, three state rows per task) - the combination none of the reduced shapes reaches."""
    fs = synthetic.make_fixed_split("ml10m", reverse_user_item_data=True, seed=0)
    rd, oms, logs, _, _, _ = _run_against_oracle(fs, H=512, B=128, aux="dropout", sparsity=[0.5, 0.5], pass_through=False,
                                                  pdrop=None, steps=2, opt="adam", l2=0.001)
    rd.close()


def test_ml10m_shape_hidden_layer_splits_the_lean_update():
    """Two layers at the ML-10M shape: the decoder rows' update runs beside the backward pass and the encoder rows'
    after it - two launches of the lean variant sharing one task list (front / back halves, one cursor each)."""
    fs = synthetic.make_fixed_split("ml10m", reverse_user_item_data=True, seed=0)
    rd, oms, logs, _, _, _ = _run_against_oracle(fs, H=512, B=128, aux="dropout", sparsity=[0.5, 0.5], pass_through=False,
                                                  pdrop=0.2, steps=3, layers=2)
    rd.close()
