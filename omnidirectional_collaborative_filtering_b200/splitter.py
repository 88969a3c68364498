"""The reference's offline splitter (`TrainValidTestSplit.py`) as a function over the native
implementation (`csrc/ocf_etl.cpp`: CSV parse, grouping, pairing, JSON/CSV writers).

The script's module-level parameters (`TrainValidTestSplit.py:17-25`) are the keyword arguments, with the
same names and defaults. The rating permutation is drawn here from the NumPy global stream exactly where
the script draws it (`np.random.permutation(num_ratings)`, `:74`), so `np.random.seed(s); split_data(...)`
writes the same files, byte for byte, as the script run after `np.random.seed(s)`.

Where the script crashes under Python 3 (np.int64 is not JSON-serialisable: all-integer CSVs, string ids
with timestamps, `save_users_and_items`), this writes what Python 2 wrote - integers as integers."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

SCHEMAS = ("movielens", "amazon", "beeradvocate", "yelp", "netflix")


def split_data(full_data_filepath="/data1/movielens/ml-1m/ratings.csv", output_filepath="data/ml1m/",
               schema_type="movielens", trainvalidtest_split=(.8, .1, .1), build_data_for_omni=True,
               include_timestamps=True, save_users_and_items=False, reverse_user_item_data=False):
    if schema_type not in SCHEMAS:
        raise ValueError("schema_type must be one of %s" % (SCHEMAS,))
    if reverse_user_item_data:                                     # TrainValidTestSplit.py:27-29
        print("Generating reverse user-item data")
        output_filepath = output_filepath + "reverse_item-user/"
    lib = _lib.lib()
    print("Loading CSV from ", full_data_filepath)
    csv = C.c_void_p()
    _lib.check(lib.ocf_csv_load(str(full_data_filepath).encode(), 3 if schema_type == "netflix" else 4, C.byref(csv)))
    try:
        n = C.c_int64()
        _lib.check(lib.ocf_csv_rows(csv, C.byref(n)))
        print("Splitting data")
        from .data_reader import sync_host_rng
        sync_host_rng()                                            # np.random may be on loan to the GPU
        order = np.ascontiguousarray(np.random.permutation(int(n.value)), dtype=np.int64)    # :74
        fractions = (C.c_double * 3)(*[float(f) for f in trainvalidtest_split])
        if os.path.dirname(output_filepath):
            os.makedirs(os.path.dirname(output_filepath), exist_ok=True)
        print("Saving splits")
        _lib.check(lib.ocf_split_write(csv, _lib.ptr(order), order.size, fractions, str(output_filepath).encode(),
                                       int(schema_type == "movielens"), int(bool(build_data_for_omni)),
                                       int(bool(include_timestamps)), int(bool(save_users_and_items)),
                                       int(bool(reverse_user_item_data))))
    finally:
        lib.ocf_csv_destroy(csv)
    return output_filepath


def split_in_memory(full_data_filepath, schema_type="movielens", trainvalidtest_split=(.8, .1, .1),
                    reverse_user_item_data=False):
    """The same split as `split_data`, straight into the stores the reader builds - no files. Draws the same one
    `np.random.permutation(num_ratings)`; returns an `ingest.LoadedSplit` for `data_reader(num_items, num_users, "",
    eval_mode="fixed_split", data=...)`. Identical to `split_data(..., save_users_and_items=True)` followed by
    `data_reader(..., use_json=True)` on its output (tests/test_split_files.py)."""
    import json
    from .ingest import LoadedSplit
    from .synthetic import Csr
    if schema_type not in SCHEMAS:
        raise ValueError("schema_type must be one of %s" % (SCHEMAS,))
    lib = _lib.lib()
    csv, split = C.c_void_p(), C.c_void_p()
    _lib.check(lib.ocf_csv_load(str(full_data_filepath).encode(), 3 if schema_type == "netflix" else 4, C.byref(csv)))
    try:
        n = C.c_int64()
        _lib.check(lib.ocf_csv_rows(csv, C.byref(n)))
        from .data_reader import sync_host_rng
        sync_host_rng()
        order = np.ascontiguousarray(np.random.permutation(int(n.value)), dtype=np.int64)    # TrainValidTestSplit.py:74
        fractions = (C.c_double * 3)(*[float(f) for f in trainvalidtest_split])
        _lib.check(lib.ocf_split_build(csv, _lib.ptr(order), order.size, fractions, int(schema_type == "movielens"),
                                       int(bool(reverse_user_item_data)), C.byref(split)))
        info = (C.c_int64 * 13)()
        _lib.check(lib.ocf_split_info(split, info))
        n_cols = int(info[0])

        def keys_of(s):
            rows, nbytes = int(info[1 + 4 * s]), int(info[2 + 4 * s])
            raw = np.empty(max(nbytes, 1), dtype=np.uint8)
            offs = np.empty(rows + 1, dtype=np.int64)
            _lib.check(lib.ocf_split_keys(split, s, _lib.ptr(raw), _lib.ptr(offs)))
            blob = raw[:nbytes].tobytes()
            return [blob[offs[k]:offs[k + 1]].decode("utf-8", "surrogatepass") for k in range(rows)]

        def store(s, part):
            rows, nnz = int(info[1 + 4 * s]), int(info[3 + 4 * s + part])
            rowptr = np.empty(rows + 1, dtype=np.int64)
            col, val = np.empty(nnz, dtype=np.int32), np.empty(nnz, dtype=np.float32)
            none = np.empty(rows, dtype=np.uint8)
            _lib.check(lib.ocf_split_csr(split, s, part, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(none)))
            return Csr(rows, n_cols, rowptr, col, val), none.astype(bool)

        need = C.c_int64()
        _lib.check(lib.ocf_split_columns_json(split, None, 0, C.byref(need)))
        text = np.empty(max(int(need.value), 1), dtype=np.uint8)
        _lib.check(lib.ocf_split_columns_json(split, _lib.ptr(text), text.size, C.byref(need)))
        unique_items = json.loads(text[:int(need.value)].tobytes().decode("ascii"))
        sets = []
        for s in (1, 2):
            ins, none = store(s, 0)
            sets.append((keys_of(s), ins, none, store(s, 1)[0]))
        return LoadedSplit(n_cols, (keys_of(0), store(0, 0)[0]), sets[0], sets[1], unique_items)
    finally:
        if split.value:
            lib.ocf_split_destroy(split)
        lib.ocf_csv_destroy(csv)
