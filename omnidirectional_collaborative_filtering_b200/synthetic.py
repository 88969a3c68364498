"""Seeded synthetic rating data in the shapes BASELINE.json names.

There is no network and the reference ships no data (`train.py:62` reads a
`datasets_metadata.json` that is not in its tree), so every workload here is
synthetic: power-law row degrees, popularity-skewed columns, unique (row, col)
pairs, ratings drawn from the dataset's rating alphabet, and the rating order
shuffled so that the per-row *stored* order is arbitrary (it matters: draw j of
the reciprocal dropout belongs to the j-th stored rating, `data_reader.py:130-134`).

The per-rating 80/10/10 split and the input/target pairing restate what
`TrainValidTestSplit.py:72-103,183-195` does:
  train   = first 80 % of a random rating permutation
  valid   = next 10 %, inputs = the train ratings of the same row (None when
            the row has no train rating)
  test    = the rest, inputs = train+valid ratings of the same row
Row keys keep first-appearance order (Python dict insertion order in the
reference), ratings inside a row keep the order of the permuted rating list.

Two products:
  * `FixedSplit`  - CSR arrays (what the B200 path consumes; scales to 1e8 ratings)
  * `to_reference_dicts()` - the JSON-shaped dicts `data_reader.py:20-80` loads,
    only sensible for small cases (tests, goldens, the CPU reference arm).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass(frozen=True)
class Shape:
    """A named dataset shape (users x items, number of ratings, rating alphabet)."""
    name: str
    n_users: int
    n_items: int
    nnz: int
    ratings: Tuple[float, ...]          # alphabet; empty => Jester-style U(-10,10) rounded to 0.01
    rating_range: float


SHAPES: Dict[str, Shape] = {
    # BASELINE.json configs[0..4]; SURVEY.md section 8(d)
    "ml1m":    Shape("ml1m", 6040, 3706, 1_000_209, (1, 2, 3, 4, 5), 4.0),
    "jester":  Shape("jester", 73_421, 100, 4_100_000, (), 20.0),
    "ml10m":   Shape("ml10m", 71_567, 10_677, 10_000_054,
                     (0.5, 1, 1.5, 2, 2.5, 3, 3.5, 4, 4.5, 5), 4.5),
    "ml20m":   Shape("ml20m", 138_493, 26_744, 20_000_263,
                     (0.5, 1, 1.5, 2, 2.5, 3, 3.5, 4, 4.5, 5), 4.5),
    "netflix": Shape("netflix", 480_189, 17_770, 100_480_507, (1, 2, 3, 4, 5), 4.0),
    # small shapes for tests / smoke
    "tiny":    Shape("tiny", 61, 47, 900, (1, 2, 3, 4, 5), 4.0),
    "small":   Shape("small", 700, 420, 30_000, (1, 2, 3, 4, 5), 4.0),
}


def _skewed_ids(rng: np.random.RandomState, n: int, size: int, alpha: float) -> np.ndarray:
    """`size` ids in [0,n) with a Zipf-like popularity profile, p(rank) ~ (rank+1)^-alpha
    (continuous inverse-CDF: rank = n * u^(1/(1-alpha))), on a shuffled id space."""
    u = rng.random_sample(size)
    ranks = np.floor(n * np.power(u, 1.0 / (1.0 - alpha))).astype(np.int64)
    np.minimum(ranks, n - 1, out=ranks)
    relabel = rng.permutation(n)
    return relabel[ranks]


def _stable_group_order(keys: np.ndarray) -> np.ndarray:
    """argsort(keys, kind='stable') for non-negative keys < 2^32, as two 16-bit radix passes
    (NumPy only uses radix sort for <= 16-bit integers; this is ~5x faster at 1e7 keys)."""
    keys = np.asarray(keys, dtype=np.int64)
    lo = (keys & 0xFFFF).astype(np.uint16)
    order = np.argsort(lo, kind="stable")
    if keys.size and int(keys.max()) >= (1 << 16):
        hi = (keys >> 16).astype(np.uint16)
        order = order[np.argsort(hi[order], kind="stable")]
    return order


def _sorted_unique(a: np.ndarray) -> np.ndarray:
    """np.unique(a) for a 1-d integer array, by sort + neighbour compare (NumPy 2.3 routes np.unique through a
    hash table that takes ~1 us per key: 2 minutes of a Netflix-sized generation)."""
    a = np.sort(a)
    if a.size == 0:
        return a
    keep = np.empty(a.size, dtype=bool)
    keep[0] = True
    np.not_equal(a[1:], a[:-1], out=keep[1:])
    return a[keep]


def make_ratings(shape: Shape, seed: int = 0, user_alpha: float = 0.45,
                 item_alpha: float = 0.65) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Unique (user, item) pairs with ratings, in a random order.

    Returns (user int32[nnz'], item int32[nnz'], rating float32[nnz']); nnz' == shape.nnz
    unless the matrix is too dense to reach it (then as many as were found).
    """
    rng = np.random.RandomState(seed)
    target = min(shape.nnz, shape.n_users * shape.n_items)
    dense_frac = target / float(shape.n_users * shape.n_items)
    if dense_frac > 0.2:
        # Jester-like: Bernoulli mask over the whole (small) matrix, per-user density varies
        dens = np.clip(rng.beta(2.0, 2.0, size=shape.n_users) * 2.0 * dense_frac, 0.02, 1.0)
        mask = rng.random_sample((shape.n_users, shape.n_items)) < dens[:, None]
        u, i = np.nonzero(mask)
        keys = u.astype(np.int64) * shape.n_items + i
    else:
        keys = np.empty(0, dtype=np.int64)
        need = target
        while need > 0:
            m = int(need * 1.25) + 1024
            u = _skewed_ids(rng, shape.n_users, m, user_alpha)
            i = _skewed_ids(rng, shape.n_items, m, item_alpha)
            keys = _sorted_unique(np.concatenate([keys, u * shape.n_items + i]))
            need = target - keys.size
        if keys.size > target:
            keys = keys[rng.permutation(keys.size)[:target]]
    keys = keys[rng.permutation(keys.size)]           # arbitrary file order
    users = (keys // shape.n_items).astype(np.int32)
    items = (keys % shape.n_items).astype(np.int32)
    if shape.ratings:
        alphabet = np.asarray(shape.ratings, dtype=np.float32)
        # mildly skewed towards the upper half of the alphabet, like real rating data
        p = np.linspace(0.6, 1.6, alphabet.size)
        p /= p.sum()
        r = alphabet[np.minimum(np.searchsorted(np.cumsum(p), rng.random_sample(keys.size), side="right"),
                                alphabet.size - 1)]
    else:
        r = np.round(rng.uniform(-10.0, 10.0, size=keys.size), 2).astype(np.float32)
    return users, items, r.astype(np.float32)


@dataclass
class Csr:
    """Row-major rating store on the host: rowptr int64[rows+1], col int32, val float32."""
    n_rows: int
    n_cols: int
    rowptr: np.ndarray
    col: np.ndarray
    val: np.ndarray

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])

    def row(self, r: int) -> Tuple[np.ndarray, np.ndarray]:
        a, b = int(self.rowptr[r]), int(self.rowptr[r + 1])
        return self.col[a:b], self.val[a:b]

    def take_rows(self, rows: np.ndarray) -> "Csr":
        """New store whose row k is a copy of this store's row rows[k] (rows[k] < 0 => empty)."""
        rows = np.asarray(rows, dtype=np.int64)
        safe = np.where(rows >= 0, rows, 0)
        lens = np.where(rows >= 0, self.rowptr[safe + 1] - self.rowptr[safe], 0)
        rowptr = np.zeros(rows.size + 1, dtype=np.int64)
        np.cumsum(lens, out=rowptr[1:])
        total = int(rowptr[-1])
        src = np.repeat(self.rowptr[safe] - rowptr[:-1], lens) + np.arange(total, dtype=np.int64)
        return Csr(rows.size, self.n_cols, rowptr, self.col[src].copy(), self.val[src].copy())


def csr_from_coo(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray, n_cols: int,
                 row_keys: Optional[np.ndarray] = None) -> Tuple[Csr, np.ndarray]:
    """Group (row, col, val) triples by row, rows in first-appearance order (dict insertion
    order in `TrainValidTestSplit.py:143-148`), ratings inside a row in arrival order.

    If `row_keys` is given it fixes the row order instead (rows absent from the data come
    out empty). Returns (csr, keys) where keys[k] is the raw row id of csr row k.
    """
    rows = np.asarray(rows, dtype=np.int64)
    if row_keys is None:
        hi0 = int(rows.max(initial=-1)) + 1
        first = np.full(hi0, rows.size, dtype=np.int64)
        first[rows[::-1]] = np.arange(rows.size - 1, -1, -1, dtype=np.int64)   # last write wins = first appearance
        present = np.flatnonzero(first < rows.size)
        row_keys = present[np.argsort(first[present], kind="stable")]
    row_keys = np.asarray(row_keys, dtype=np.int64)
    hi = int(max(rows.max(initial=-1), row_keys.max(initial=-1))) + 1
    rank = np.full(hi, -1, dtype=np.int64)
    rank[row_keys] = np.arange(row_keys.size)
    rr = rank[rows]
    keep = rr >= 0
    if not keep.all():
        rr, cols, vals = rr[keep], cols[keep], vals[keep]
    order = _stable_group_order(rr)
    counts = np.bincount(rr, minlength=row_keys.size)
    rowptr = np.zeros(row_keys.size + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    csr = Csr(row_keys.size, n_cols, rowptr, np.ascontiguousarray(cols[order], dtype=np.int32),
              np.ascontiguousarray(vals[order], dtype=np.float32))
    return csr, row_keys


@dataclass
class FixedSplit:
    """Everything `eval_mode="fixed_split"` needs, as CSR stores with aligned row indexing.

    train            rows = train_keys (dict order of ratingsBy*_dicts_train)
    valid_in/valid_tg rows = valid_keys; valid_in row k = train ratings of valid_keys[k]
                      (`none` flag where the reference stores None)
    test_in/test_tg   rows = test_keys; inputs = train+valid ratings
    """
    n_cols: int
    train: Csr
    train_keys: np.ndarray
    valid_in: Csr
    valid_tg: Csr
    valid_keys: np.ndarray
    valid_none: np.ndarray
    test_in: Csr
    test_tg: Csr
    test_keys: np.ndarray
    test_none: np.ndarray
    rating_range: float = 4.0
    meta: dict = field(default_factory=dict)


def build_fixed_split(users: np.ndarray, items: np.ndarray, ratings: np.ndarray,
                      n_users: int, n_items: int, reverse_user_item_data: bool,
                      seed: int = 1, fractions=(0.8, 0.1, 0.1),
                      rating_range: float = 4.0) -> FixedSplit:
    """Per-rating split + pairing, `TrainValidTestSplit.py:72-103` restated on arrays.

    With `reverse_user_item_data` rows are items and columns users (`train.py:71-76`).
    Column ids are already dense (0..n_cols-1).
    """
    rng = np.random.RandomState(seed)
    n = users.size
    order = rng.permutation(n)                         # :74
    n_tr = int(n * fractions[0])                       # :76-78
    n_va = int(n * fractions[1])
    tr, va, te = order[:n_tr], order[n_tr:n_tr + n_va], order[n_tr + n_va:]
    te_in = order[:n_tr + n_va]                        # :83
    if reverse_user_item_data:
        rows, cols, n_cols = items, users, n_users
    else:
        rows, cols, n_cols = users, items, n_items

    def group(idx, keys=None):
        return csr_from_coo(rows[idx], cols[idx], ratings[idx], n_cols, keys)

    train, train_keys = group(tr)
    valid_tg, valid_keys = group(va)
    test_tg, test_keys = group(te)

    def paired_inputs(src_idx, keys):
        # map_inputs_to_targets (:183-195): the input row is the whole row of the input set
        src, src_keys = group(src_idx)
        hi = int(max(src_keys.max(initial=-1), keys.max(initial=-1))) + 1
        pos = np.full(hi, -1, dtype=np.int64)
        pos[src_keys] = np.arange(src_keys.size)
        where = pos[keys]
        return src.take_rows(where), (where < 0)

    valid_in, valid_none = paired_inputs(tr, valid_keys)
    test_in, test_none = paired_inputs(te_in, test_keys)
    return FixedSplit(n_cols, train, train_keys, valid_in, valid_tg, valid_keys, valid_none,
                      test_in, test_tg, test_keys, test_none, rating_range,
                      {"reverse_user_item_data": reverse_user_item_data, "seed": seed})


def make_fixed_split(shape_name: str, reverse_user_item_data: bool = True, seed: int = 0) -> FixedSplit:
    shape = SHAPES[shape_name]
    u, i, r = make_ratings(shape, seed)
    return build_fixed_split(u, i, r, shape.n_users, shape.n_items, reverse_user_item_data,
                             seed + 1, rating_range=shape.rating_range)


# ---------------------------------------------------------------------------------------
# Reference-shaped dicts (small cases only)
# ---------------------------------------------------------------------------------------

def _row_lists(csr: Csr, col_ids: List, none: Optional[np.ndarray] = None) -> List[Optional[list]]:
    out: List[Optional[list]] = []
    for r in range(csr.n_rows):
        if none is not None and none[r]:
            out.append(None)
            continue
        c, v = csr.row(r)
        out.append([[col_ids[int(ci)], float(vi)] for ci, vi in zip(c, v)])
    return out


def to_reference_dicts(fs: FixedSplit, raw_col_id=lambda c: 3 * c + 7,
                       raw_row_key=lambda k: str(int(k))) -> dict:
    """The in-memory equivalent of the JSON files `data_reader.py:20-70` loads.

    Keys are strings (JSON object keys; `TrainValidTestSplit.py:127`), column ids are raw
    (non-dense) so the id->dense map (`data_reader.py:24-28`) is exercised.
    Returns {"unique_cols": [...], "train": {...}, "valid": [in, tg], "test": [in, tg]}.
    """
    col_ids = [raw_col_id(c) for c in range(fs.n_cols)]

    def as_dict(keys, lists):
        return {raw_row_key(k): l for k, l in zip(keys, lists)}

    return {
        "unique_cols": col_ids,
        "train": as_dict(fs.train_keys, _row_lists(fs.train, col_ids)),
        "valid": [as_dict(fs.valid_keys, _row_lists(fs.valid_in, col_ids, fs.valid_none)),
                  as_dict(fs.valid_keys, _row_lists(fs.valid_tg, col_ids))],
        "test": [as_dict(fs.test_keys, _row_lists(fs.test_in, col_ids, fs.test_none)),
                 as_dict(fs.test_keys, _row_lists(fs.test_tg, col_ids))],
    }
