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


# Multi-tile shapes: the persistent kernel gives every CTA pair several (256-column x row-chunk) tiles, so the
# TMEM accumulator ping-pong (tfull / tempty phases), the smem ring wrapping across tiles and the tile scheduler all
# run for more than one round: 10 677 columns x 1 024 rows = 42 x 4 = 168 pair tiles, 71 567 x 1 024 (bench.py's
# scoring shape) = 280 x 4 = 1 120 pair tiles, over 74 CTA pairs.
@pytest.mark.parametrize("reverse", [False, True], ids=["ml10m_users_168_tiles", "ml10m_items_1120_tiles"])
def test_scores_match_oracle_on_many_tiles_per_cta_pair(reverse):
    from tests.helpers import cached_split
    fs = cached_split("ml10m", reverse)
    rows = 1024
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    assert rd.val_set_size >= rows
    om, ref = _models(fs.n_cols, 1, 512, rows, "sigmoid", "dropout", seed=5)
    n_pair_tiles = ((fs.n_cols + 255) // 256) * ((rows + 255) // 256)
    assert n_pair_tiles >= 2 * 74                      # every CTA pair of a 148-SM part gets at least two tiles
    np.random.seed(9)
    batch = next(rd.data_gen(rows, None, "valid", True, "dropout", -1))
    got = om.model.score(batch).astype(np.float64)
    again = om.model.score(batch)
    feed, targets = batch
    _, want, acts, _, _ = ref.forward(feed)
    assert got.shape == want.shape == (rows, fs.n_cols)
    assert np.array_equal(got.astype(np.float32), again)          # same bits on a second call
    scale = np.abs(acts[-1]) @ np.abs(ref.get_weights()[-2]) + 1.0
    err = got - want
    assert np.max(np.abs(err) / scale) <= 2e-3
    assert abs(err.mean()) <= 2e-5 * max(1.0, np.abs(want).mean())
    # every 256-column tile and every row chunk carries its own data (a stale or swapped accumulator stage would
    # put one tile's scores under another tile's columns): per-tile RMS error stays at tf32 level everywhere
    ncol = fs.n_cols // 256 * 256
    tile_rms = np.sqrt((err[:, :ncol] ** 2).reshape(rows // 256, 256, ncol // 256, 256).mean(axis=(1, 3)))
    ref_rms = np.sqrt((want[:, :ncol] ** 2).reshape(rows // 256, 256, ncol // 256, 256).mean(axis=(1, 3)))
    assert np.all(tile_rms <= 2e-3 * (ref_rms + 1.0))
    mask, obs = feed[-1], targets != 0
    rm_got = np.sqrt((((mask * got) - targets)[obs] ** 2).mean())
    rm_want = np.sqrt((((mask * want) - targets)[obs] ** 2).mean())
    assert abs(rm_got - rm_want) <= 1e-3 * rm_want
    # top-k over the same multi-tile scores: exact selection on the device's own scores
    cols, scores = om.model.recommend(batch, k=50, exclude_seen=False)
    order = np.lexsort((np.arange(fs.n_cols)[None, :].repeat(8, 0), -again[:8]), axis=1)[:, :50]
    assert np.array_equal(cols[:8], order.astype(np.int32))
    assert np.array_equal(scores[:8], np.take_along_axis(again[:8], order, axis=1))
    om.model.close()
    rd.close()
