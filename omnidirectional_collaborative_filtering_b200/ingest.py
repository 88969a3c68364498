"""The reference's on-disk JSON files -> CSR arrays through the native parser (`csrc/ocf_etl.cpp`,
`ocf_vocab_*` / `ocf_ratings_*` in include/ocf.h). Replaces `json.load` + the per-rating dict lookups of
`data_reader.py:85-92,134-136` at load time: no Python object per rating, so a Netflix-sized file
(10^8 ratings) loads in seconds within a few GB instead of tens of GB of lists."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .synthetic import Csr


class LoadedSplit(object):
    """What `eval_mode="fixed_split"` needs, as keys + CSR stores: train = (keys, csr); valid / test = (keys, input
    csr, none flags, target csr) with aligned rows. Produced by the native file parser (`load_split`) or straight
    from a ratings CSV (`splitter.split_in_memory`); `data_reader(..., data=LoadedSplit)` consumes it."""

    def __init__(self, n_cols, train, valid, test, unique_items=None):
        self.n_cols, self.train, self.valid, self.test, self.unique_items = int(n_cols), train, valid, test, unique_items


class Vocab(object):
    """`unique_items_list.json` / `unique_users_list.json`: id -> dense column (data_reader.py:24-28)."""

    def __init__(self, path):
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().ocf_vocab_load_json(str(path).encode(), C.byref(self.handle)))
        n = C.c_int64()
        _lib.check(_lib.lib().ocf_vocab_size(self.handle, C.byref(n)))
        self.size = int(n.value)

    def close(self):
        if self.handle is not None and self.handle.value:
            _lib.lib().ocf_vocab_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_ratings(path, vocab: Vocab, paired: bool, n_cols=None):
    """One rating-dict file. Returns (keys, csr) for a single dict, (keys, inputs csr, none flags,
    targets csr) for an [inputs, targets] pair; rows follow `keys` (see ocf_ratings_load_json)."""
    lib = _lib.lib()
    h = C.c_void_p()
    _lib.check(lib.ocf_ratings_load_json(str(path).encode(), vocab.handle, int(bool(paired)), C.byref(h)))
    try:
        info = (C.c_int64 * 4)()
        _lib.check(lib.ocf_ratings_info(h, info))
        rows, key_bytes = int(info[0]), int(info[1])
        raw = np.empty(max(key_bytes, 1), dtype=np.uint8)
        offs = np.empty(rows + 1, dtype=np.int64)
        _lib.check(lib.ocf_ratings_keys(h, _lib.ptr(raw), _lib.ptr(offs)))
        blob = raw[:key_bytes].tobytes()
        keys = [blob[offs[k]:offs[k + 1]].decode("utf-8", "surrogatepass") for k in range(rows)]
        n_cols = vocab.size if n_cols is None else int(n_cols)

        def store(which):
            nnz = int(info[2 + which])
            rowptr = np.empty(rows + 1, dtype=np.int64)
            col = np.empty(nnz, dtype=np.int32)
            val = np.empty(nnz, dtype=np.float32)
            none = np.empty(rows, dtype=np.uint8)
            _lib.check(lib.ocf_ratings_csr(h, which, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(none)))
            return Csr(rows, n_cols, rowptr, col, val), none.astype(bool)

        if not paired:
            return keys, store(0)[0]
        ins, none = store(0)
        return keys, ins, none, store(1)[0]
    finally:
        lib.ocf_ratings_destroy(h)
