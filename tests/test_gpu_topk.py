"""GPU parity of the top-k serving epilogue (`ocf_score_topk`: sample pivot + radix select + bitonic
sort, one CTA per row) against the NumPy oracle applied to the scores the device itself produced
(`model.score`): same winners, same order, same values - bit-exact, ties included."""
import numpy as np
import pytest

from oracle import ref_topk
from omnidirectional_collaborative_filtering_b200 import synthetic
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
from omnidirectional_collaborative_filtering_b200.model import omni_model

pytestmark = pytest.mark.gpu


def _seen(batch):
    csr = batch.source.in_store.csr
    return [csr.col[csr.rowptr[r]:csr.rowptr[r + 1]] for r in batch.rows]


def _setup(shape, rev, B, width, seed=3, flat=False):
    fs = synthetic.make_fixed_split(shape, reverse_user_item_data=rev, seed=seed)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    np.random.seed(seed)
    om = omni_model(1, width, fs.n_cols, B, dense_activation="sigmoid", use_causal_info=False)
    w = om.model.get_weights()
    rs = np.random.RandomState(seed + 1)
    if flat:                                        # scores = bias only: a handful of distinct values, thousands of exact ties
        w[-2][:] = 0.0
        w[-1] = rs.randint(0, 5, size=w[-1].shape).astype(np.float32)
    else:
        w[-2] = (rs.normal(size=w[-2].shape) * 0.3).astype(np.float32)
        w[-1] = rs.normal(size=w[-1].shape).astype(np.float32)
    om.model.set_weights(w)
    batch = next(rd.data_gen(min(B, rd.val_set_size), None, "valid", True, None, -1))
    return rd, om, batch


@pytest.mark.parametrize("shape,rev,B,width,k,exclude", [
    ("tiny", True, 8, 12, 5, True),                 # k close to the catalogue width
    ("tiny", False, 8, 12, 47, False),              # k == n_cols
    ("small", True, 64, 100, 10, True),             # 700 columns: everything is a candidate
    ("ml1m", True, 96, 64, 100, True),              # 6040 columns: sampled pivot
    ("ml1m", True, 40, 64, 512, False),             # largest k
])
def test_topk_matches_oracle_on_device_scores(shape, rev, B, width, k, exclude):
    rd, om, batch = _setup(shape, rev, B, width)
    scores = om.model.score(batch)
    got_c, got_v = om.model.recommend(batch, k=k, exclude_seen=exclude)
    want_c, want_v = ref_topk.topk(scores, k, _seen(batch) if exclude else None)
    assert np.array_equal(got_c, want_c)
    assert np.array_equal(got_v, want_v)
    if exclude:
        for b, cols in enumerate(_seen(batch)):
            assert not set(got_c[b][got_c[b] >= 0].tolist()) & set(cols.tolist())
    rd.close()


@pytest.mark.parametrize("k", [7, 100])
def test_topk_with_massive_ties(k):
    """Bias-only scores: 5 distinct values over 6040 columns. The sample pivot admits every column,
    the select falls back to the whole row and the ties are taken in column order."""
    rd, om, batch = _setup("ml1m", True, 16, 32, flat=True)
    scores = om.model.score(batch)
    got_c, got_v = om.model.recommend(batch, k=k, exclude_seen=True)
    want_c, want_v = ref_topk.topk(scores, k, _seen(batch))
    assert np.array_equal(got_c, want_c) and np.array_equal(got_v, want_v)
    rd.close()
