"""The synthetic generator's split/pairing against a naive dict-based restatement of
TrainValidTestSplit.py:72-103,121-149,183-195."""
import numpy as np

from omnidirectional_collaborative_filtering_b200 import synthetic as S


def _naive(users, items, ratings, reverse, seed, fractions=(0.8, 0.1, 0.1)):
    rng = np.random.RandomState(seed)
    n = users.size
    order = rng.permutation(n)
    n_tr, n_va = int(n * fractions[0]), int(n * fractions[1])
    rows, cols = (items, users) if reverse else (users, items)

    def build(idx):
        d = {}
        for k in idx:
            d.setdefault(int(rows[k]), []).append((int(cols[k]), float(ratings[k])))
        return d

    train, valid, test = build(order[:n_tr]), build(order[n_tr:n_tr + n_va]), build(order[n_tr + n_va:])
    test_in = build(order[:n_tr + n_va])
    pair = lambda src, tgt: {k: src.get(k) for k in tgt}
    return train, (pair(train, valid), valid), (pair(test_in, test), test)


def _as_dict(csr, keys, none=None):
    out = {}
    for r, k in enumerate(keys):
        if none is not None and none[r]:
            out[int(k)] = None
            continue
        c, v = csr.row(r)
        out[int(k)] = [(int(ci), float(vi)) for ci, vi in zip(c, v)]
    return out


def test_fixed_split_matches_naive_dict_build():
    for reverse in (True, False):
        sh = S.SHAPES["small"]
        u, i, r = S.make_ratings(sh, 3)
        assert u.size == sh.nnz and np.unique(u.astype(np.int64) * sh.n_items + i).size == sh.nnz
        fs = S.build_fixed_split(u, i, r, sh.n_users, sh.n_items, reverse, seed=9)
        train, valid, test = _naive(u, i, r, reverse, 9)
        got = _as_dict(fs.train, fs.train_keys)
        assert list(got.keys()) == list(train.keys()) and got == train          # key order = dict insertion order
        assert _as_dict(fs.valid_tg, fs.valid_keys) == valid[1]
        assert list(_as_dict(fs.valid_tg, fs.valid_keys)) == list(valid[1])
        assert _as_dict(fs.valid_in, fs.valid_keys, fs.valid_none) == valid[0]
        assert _as_dict(fs.test_tg, fs.test_keys) == test[1]
        assert _as_dict(fs.test_in, fs.test_keys, fs.test_none) == test[0]


def test_stable_group_order_is_stable_argsort():
    rs = np.random.RandomState(0)
    for hi in (10, 70000, 200000):
        k = rs.randint(0, hi, size=50000)
        assert np.array_equal(S._stable_group_order(k), np.argsort(k, kind="stable"))


def test_reference_dicts_round_trip():
    fs = S.make_fixed_split("tiny", True, seed=2)
    d = S.to_reference_dicts(fs)
    assert len(d["train"]) == fs.train.n_rows and len(d["valid"][1]) == fs.valid_tg.n_rows
    k0 = str(int(fs.train_keys[0]))
    c, v = fs.train.row(0)
    assert d["train"][k0] == [[3 * int(ci) + 7, float(vi)] for ci, vi in zip(c, v)]
