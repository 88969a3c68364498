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
