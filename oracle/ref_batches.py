"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's batch
construction, `/root/reference/data_reader.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module. Nothing under `omnidirectional_collaborative_filtering_b200/`
does.

Parity status: PINNED. `tests/golden/make_golden.py` imports the reference's own
`data_reader.py` (unmodified, behind a stub `tensorflow` module) in the build container and
stores its outputs under `tests/golden/*.npz`; `tests/test_oracle_batches.py` checks every
function here against those files bit for bit, including the position of the NumPy global
RNG stream after the run.

What is restated (reference file:line):
  RefData                 - ctor state, `data_reader.py:12-80` (id->dense map :24-28, row maps
                            :30-44, set sizes/orders :72-80)
  split_rows              - `split_for_validation`, `data_reader.py:300-312`
  split_batch_loop        - `build_sparse_batch`, dense branch, `data_reader.py:95-200`
                            (per-rating Python loop: this IS the reference's cost model and is
                            what the CPU baseline times)
  split_batch_vec         - same result and same RNG consumption, vectorised (fast checker)
  fixed_batch_loop        - `build_sparse_batch_fixed_split`, dense branch, `:202-298`
  feed_list               - aux-mask selection / input list assembly, `:341-361`, `:389-411`
  batch_stream            - `data_gen`, `:314-419`

Not restated: timestamps (`useTimestamps`, broken in the reference at :132/:359/:409) and the
scipy COO `sparse_representation` branch (needs a patched Keras backend, `train.py:53`).
"""
from __future__ import annotations

import numpy as np

AUX_TYPES = ("causal", "dropout", "zeros", "both", None)


class RefData(object):
    """State the reference reader keeps after its constructor ran."""

    def __init__(self, num_items, num_users, unique_cols, eval_mode="ablation",
                 user_dict=None, train=None, valid=None, test=None,
                 nonsequentialusers=False, unique_rows=None):
        self.num_items = int(num_items)          # width N of every batch array (:13)
        self.num_users = int(num_users)
        self.eval_mode = eval_mode
        # :24-28 column id -> dense position
        self.col_of = {}
        for pos, raw in enumerate(unique_cols):
            self.col_of[raw] = pos
        # :30-44 dense row -> raw row id (ablation mode only uses it, :125)
        if nonsequentialusers:
            self.row_of_dense = {pos: raw for pos, raw in enumerate(unique_rows)}
        else:
            self.row_of_dense = {pos: pos for pos in range(self.num_users)}
        if eval_mode == "ablation":
            self.user_dict = user_dict           # :55
        elif eval_mode == "fixed_split":
            self.user_dict = train               # :67-68
            self.valid = valid                   # (input dict, target dict) :69
            self.test = test                     # :70
            self.train_set = list(train.keys())              # :78
            self.val_set = list(valid[1].keys())             # :79
            self.test_set = list(test[1].keys())             # :80
            self.train_set_size = len(self.train_set)        # :73
            self.val_set_size = len(self.val_set)            # :74
            self.test_set_size = len(self.test_set)          # :75
        else:
            raise ValueError("eval_mode must be 'ablation' or 'fixed_split'")


def split_rows(data, val_split, seed=None, rng=np.random):
    """`split_for_validation` (:300-312): one permutation of the dense row ids, cut 3 ways."""
    if seed is not None:
        rng.seed(seed)
    order = rng.permutation(data.num_users)
    n_tr = int(data.num_users * val_split[0])
    n_va = int(data.num_users * val_split[1])
    data.train_set_size = n_tr
    data.val_set_size = n_va
    data.test_set_size = int(data.num_users * val_split[2])
    data.train_set = order[:n_tr]
    data.val_set = order[n_tr:n_tr + n_va]
    data.test_set = order[n_tr + n_va:]


def _ratings_of(data, key):
    if data.eval_mode == "ablation":
        key = data.row_of_dense[key]             # :124-125
    return data.user_dict[key]                   # :128


def split_batch_loop(data, order, batch_size, start, sparsity, aux_value,
                     pass_through=False, rng=np.random):
    """Reciprocal random input/target split of B rows, one rating at a time (:109-170).

    Returns (mask_in, mask_out, x, t, observed), each float64 [B, N].
    """
    B, N = int(batch_size), data.num_items
    mask_in = np.zeros([B, N])
    x = np.zeros([B, N])
    observed = np.zeros([B, N])
    t = np.zeros([B, N])
    mask_out = np.zeros([B, N])
    keep_prob = rng.uniform(low=sparsity[0], high=sparsity[1], size=B)      # :120
    for b in range(B):
        lst = _ratings_of(data, order[start + b])
        s = keep_prob[b]
        as_input = rng.choice([0, 1], size=len(lst), p=[1 - s, s])          # :130
        for j, pair in enumerate(lst):
            c = data.col_of[pair[0]]                                         # :135
            r = pair[1]
            if as_input[j] == 1:                                             # :158-163
                mask_in[b, c] = aux_value
                x[b, c] = r
                if pass_through:
                    mask_out[b, c] = aux_value
                    t[b, c] = r
            else:                                                            # :164-166
                mask_out[b, c] = aux_value
                t[b, c] = r
            observed[b, c] = aux_value                                       # :169
    return mask_in, mask_out, x, t, observed


def draw_split_flags(lengths, sparsity, rng=np.random):
    """The RNG draws of one batch, vectorised: returns (keep_prob[B], flags uint8[sum n]).

    `choice([0,1], n, p=[1-s, s])` draws n doubles u and returns 1 where u >= cdf0 with
    cdf = cumsum(p) / cumsum(p)[-1]; B consecutive calls read one contiguous run of the
    stream, so a single random_sample(sum n) is the same numbers.
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    keep_prob = rng.uniform(low=sparsity[0], high=sparsity[1], size=lengths.size)
    u = rng.random_sample(int(lengths.sum()))
    p0 = 1 - keep_prob
    cdf0 = p0 / (p0 + keep_prob)
    flags = (u >= np.repeat(cdf0, lengths)).astype(np.uint8)
    return keep_prob, flags


def split_batch_vec(data, order, batch_size, start, sparsity, aux_value,
                    pass_through=False, rng=np.random):
    """Same arrays and the same RNG consumption as `split_batch_loop`, without the
    per-rating Python loop (duplicates inside a row still resolve last-write-wins)."""
    B, N = int(batch_size), data.num_items
    rows = [_ratings_of(data, order[start + b]) for b in range(B)]
    lengths = np.array([len(l) for l in rows], dtype=np.int64)
    _, flags = draw_split_flags(lengths, sparsity, rng)
    out = [np.zeros([B, N]) for _ in range(5)]
    mask_in, mask_out, x, t, observed = out
    pos = 0
    for b, lst in enumerate(rows):
        n = len(lst)
        if n == 0:
            continue
        cols = np.fromiter((data.col_of[p[0]] for p in lst), dtype=np.int64, count=n)
        vals = np.fromiter((p[1] for p in lst), dtype=np.float64, count=n)
        f = flags[pos:pos + n].astype(bool)
        pos += n
        # fancy assignment applies in index order, so repeated columns keep the last write
        mask_in[b, cols[f]] = aux_value
        x[b, cols[f]] = vals[f]
        tg = np.ones(n, dtype=bool) if pass_through else ~f
        mask_out[b, cols[tg]] = aux_value
        t[b, cols[tg]] = vals[tg]
        observed[b, cols] = aux_value
    return mask_in, mask_out, x, t, observed


def fixed_batch_loop(data, in_dict, tgt_dict, order, batch_size, start, aux_value):
    """Valid/test batch from paired dicts (:215-268). Returns the five arrays + target_count."""
    B, N = int(batch_size), data.num_items
    mask_in = np.zeros([B, N])
    x = np.zeros([B, N])
    observed = np.zeros([B, N])
    t = np.zeros([B, N])
    mask_out = np.zeros([B, N])
    n_targets = 0
    for b in range(B):
        key = order[start + b]
        given = in_dict[key]
        wanted = tgt_dict[key]
        if given is not None:                                                # :234-252
            for pair in given:
                c = data.col_of[pair[0]]
                mask_in[b, c] = aux_value
                x[b, c] = pair[1]
                observed[b, c] = aux_value
        for pair in wanted:                                                  # :256-268
            c = data.col_of[pair[0]]
            mask_out[b, c] = aux_value
            t[b, c] = pair[1]
            observed[b, c] = aux_value
            n_targets += 1
    return mask_in, mask_out, x, t, observed, n_targets


def feed_list(arrays, aux_type):
    """What the generator hands to Keras (:341-361): [x, (aux), mask_out, (observed)]."""
    mask_in, mask_out, x, t, observed = arrays[:5]
    if aux_type is None:
        return [x, mask_out]
    if aux_type == "causal":
        aux = observed
    elif aux_type in ("dropout", "both"):
        aux = mask_in
    elif aux_type == "zeros":
        aux = np.zeros_like(mask_in)
    else:
        raise ValueError("unknown auxilliary_mask_type %r" % (aux_type,))
    feed = [x, aux, mask_out]
    if aux_type == "both":
        feed.append(observed)
    return feed


def batch_stream(data, batch_size, data_sparsity, train_val_test="train", shuffle=True,
                 auxilliary_mask_type="dropout", aux_var_value=-1, return_target_count=False,
                 pass_through_input_training=False, rng=np.random, vectorised=False):
    """`data_gen` (:314-419): lazy permutation, floor(n/B) batches, then None for ever."""
    if train_val_test == "train":
        order, n = data.train_set, data.train_set_size
    elif train_val_test == "valid":
        order, n = data.val_set, data.val_set_size
    elif train_val_test == "test":
        order, n = data.test_set, data.test_set_size
    else:
        raise ValueError(train_val_test)
    if shuffle:
        order = rng.permutation(order)                                       # :326-327
    n_batches = int(np.floor(n / batch_size))                                # :329
    build = split_batch_vec if vectorised else split_batch_loop
    if data.eval_mode == "ablation" or train_val_test == "train":            # :331
        for i in range(n_batches):
            arrays = build(data, order, batch_size, i * batch_size, data_sparsity,
                           aux_var_value, pass_through_input_training, rng)
            yield (feed_list(arrays, auxilliary_mask_type), arrays[3])
    else:                                                                    # :366
        in_dict, tgt_dict = data.valid if train_val_test == "valid" else data.test
        for i in range(n_batches):
            arrays = fixed_batch_loop(data, in_dict, tgt_dict, order, batch_size,
                                      i * batch_size, aux_var_value)
            feed = feed_list(arrays, auxilliary_mask_type)
            if return_target_count:
                yield (feed, arrays[3], arrays[5])
            else:
                yield (feed, arrays[3])
    while True:                                                              # :418-419
        yield None
