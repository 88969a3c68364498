"""GPU parity of full-catalogue scoring (`ocf_score`: tcgen05 kind::tf32 GEMM, accumulators in
TMEM) against the fp64 oracle. The operands are rounded to tf32 (10 mantissa bits, unit
round-off 2^-11) on their way into shared memory; the products accumulate in fp32. Tolerances:
  * per score: |err| <= 2e-3 * (|h|.|w| + 1), far inside what 2^-11 per operand allows;
  * no systematic shrink: |mean signed error| <= 2e-5 (truncation instead of rounding fails this);
  * RMSE computed from the scores within 1e-3 relative of the oracle's (the north-star bar)."""
import numpy as np
import pytest

from oracle import ref_model
from omnidirectional_collaborative_filtering_b200 import synthetic
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
from omnidirectional_collaborative_filtering_b200.model import omni_model

pytestmark = pytest.mark.gpu


def _models(N, layers, width, B, act, aux, seed):
    kw = dict(dense_activation=act, use_causal_info=aux is not None)
    np.random.seed(seed)
    om = omni_model(layers, width, N, B, auxilliary_mask_type=aux, **kw)
    ref = ref_model.RefModel(layers, width, N, B, dtype=np.float64, rng=np.random.RandomState(0), **kw)
    w = om.model.get_weights()
    rs = np.random.RandomState(seed + 1)
    for i in range(1, len(w), 2):
        w[i] = (rs.normal(size=w[i].shape) * 0.1).astype(np.float32)
    w[-2] = (rs.normal(size=w[-2].shape) * 0.3).astype(np.float32)          # decoder kernel with some weight
    om.model.set_weights(w)
    ref.set_weights([x.astype(np.float64) for x in w])
    return om, ref


# (shape, reverse, layers, width, B, aux): N not a multiple of 128, widths that pad to 128..512,
# batches below / at / above the 64-128-256 row chunks incl. ragged last chunks
CASES = [
    ("tiny", True, 1, 12, 8, None),
    ("small", True, 1, 200, 64, None),
    ("small", False, 1, 500, 128, "dropout"),
    ("small", True, 2, 300, 300, None),
    ("small", False, 1, 100, 400, "causal"),
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_scores_match_oracle(case):
    shape, rev, layers, width, B, aux = case
    fs = synthetic.make_fixed_split(shape, reverse_user_item_data=rev, seed=11)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    om, ref = _models(fs.n_cols, layers, width, B, "sigmoid", aux, seed=5)
    np.random.seed(9)
    B = min(B, rd.val_set_size)
    gen = rd.data_gen(B, None, "valid", True, aux, -1)
    for _ in range(2):
        batch = next(gen)
        if batch is None:
            break
        got = om.model.score(batch).astype(np.float64)
        feed, targets = batch                      # the reference's dense arrays (CUDA scatter)
        _, want, acts, _, _ = ref.forward(feed)
        assert got.shape == want.shape == (B, fs.n_cols)
        scale = np.abs(acts[-1]) @ np.abs(ref.get_weights()[-2]) + 1.0
        err = got - want
        assert np.max(np.abs(err) / scale) <= 2e-3
        assert abs(err.mean()) <= 2e-5 * max(1.0, np.abs(want).mean())
        # masked RMSE from the scores (train.py:243-252 with y = mask * full)
        mask = feed[-1] if aux != "both" else feed[-2]
        obs = targets != 0
        if obs.any():
            rm_got = np.sqrt((((mask * got) - targets)[obs] ** 2).mean())
            rm_want = np.sqrt((((mask * want) - targets)[obs] ** 2).mean())
            assert abs(rm_got - rm_want) <= 1e-3 * rm_want
    rd.close()


def test_score_rows_are_independent():
    """Scoring a row alone or inside a 256-row chunk gives the same bits (one TMEM column per row)."""
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=3)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    om, _ = _models(fs.n_cols, 1, 256, 256, "tanh", None, seed=2)
    B = min(256, rd.val_set_size)
    big = next(rd.data_gen(B, None, "valid", False, None, -1))
    small = next(rd.data_gen(8, None, "valid", False, None, -1))
    a = om.model.score(big)
    b = om.model.score(small)
    assert np.array_equal(a[:8], b)
    c = om.model.score(big, reuse_output=True)          # page-locked destination owned by the model
    assert np.array_equal(a, c)
    rd.close()
