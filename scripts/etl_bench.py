#!/usr/bin/env python
"""Host-side throughput of the two file-format neighbours (no GPU involved): the native splitter against
the reference's `TrainValidTestSplit.py` (run unmodified on a bounded sample when /root/reference exists),
and the native JSON ingest against `json.load` + the per-rating Python loop, on an ML-1M-shaped CSV.

    python scripts/etl_bench.py [--ratings 1000209] [--ref-sample 20000] > profiles/rNN_etl_cpu.json
"""
import argparse
import contextlib
import io
import json
import os
import re
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnidirectional_collaborative_filtering_b200 import ingest, splitter, synthetic
from omnidirectional_collaborative_filtering_b200.data_reader import _csr_from_lists

REF = "/root/reference/TrainValidTestSplit.py"


def write_csv(path, u, i, r, n):
    with open(path, "w") as f:
        f.write("userId,movieId,rating,timestamp\n")
        ts = 978300000 + (np.arange(n) * 7919) % 10 ** 6
        for k in range(n):
            f.write("%d,%d,%s,%d\n" % (u[k] + 1, i[k] + 1, repr(float(r[k])), ts[k]))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ratings", type=int, default=1000209)
    ap.add_argument("--ref-sample", type=int, default=20000)
    args = ap.parse_args()
    shape = synthetic.SHAPES["ml1m"]
    u, i, r = synthetic.make_ratings(shape, 0)
    n = min(args.ratings, u.size)
    d = tempfile.mkdtemp(prefix="ocf_etl_") + "/"
    write_csv(d + "ratings.csv", u, i, r, n)
    out = {"workload": "ml1m-shaped synthetic CSV, %d ratings, schema movielens, include_timestamps=False" % n,
           "cores": {"native_splitter": "1 thread for the CSV parse and the user index, then one thread per output file (5)",
                     "reference_splitter": 1, "native_ingest": "1 (files one after the other); the reader parses its 3 files on 3 threads",
                     "python_ingest": 1}}
    np.random.seed(1)
    t0 = time.perf_counter()
    quiet(splitter.split_data, d + "ratings.csv", d + "native/", "movielens", include_timestamps=False,
          save_users_and_items=True)
    t = time.perf_counter() - t0
    out["native_splitter"] = {"seconds": t, "ratings_per_s": n / t}
    if os.path.exists(REF):
        m = min(args.ref_sample, n)
        write_csv(d + "sample.csv", u, i, r, m)
        os.makedirs(d + "ref/")
        src = open(REF).read()
        for k, v in dict(full_data_filepath=d + "sample.csv", output_filepath=d + "ref/", schema_type="movielens",
                         include_timestamps=False, save_users_and_items=False, reverse_user_item_data=False).items():
            src = re.sub(r"(?m)^%s = .*$" % k, "%s = %r" % (k, v), src, count=1)
        np.random.seed(1)
        t0 = time.perf_counter()
        quiet(exec, compile(src, "TrainValidTestSplit.py", "exec"), {"__name__": "reference_split"})
        t = time.perf_counter() - t0
        out["reference_splitter"] = {"seconds": t, "ratings_per_s": m / t, "sample": "%d ratings of the same CSV" % m}
        out["splitter_speedup"] = out["native_splitter"]["ratings_per_s"] / out["reference_splitter"]["ratings_per_s"]
    files = [("ratingsByUser_dicts_train", False), ("ratingsByUser_dicts_valid", True), ("ratingsByUser_dicts_test", True)]
    t0 = time.perf_counter()
    vocab = ingest.Vocab(d + "native/unique_items_list.json")
    total = 0
    for name, paired in files:
        got = ingest.load_ratings(d + "native/" + name + ".json", vocab, paired)
        total += got[1].nnz + (got[3].nnz if paired else 0)
    t = time.perf_counter() - t0
    mb = sum(os.path.getsize(d + "native/" + name + ".json") for name, _ in files) / 1e6
    out["native_ingest"] = {"seconds": t, "ratings_per_s": total / t, "MB_per_s": mb / t, "ratings": total}
    t0 = time.perf_counter()
    with open(d + "native/unique_items_list.json") as f:
        ids = json.load(f)
    col_of = {x: k for k, x in enumerate(ids)}
    for name, paired in files:
        with open(d + "native/" + name + ".json") as f:
            obj = json.load(f)
        if paired:
            keys = list(obj[1].keys())
            _csr_from_lists([obj[0][k] for k in keys], col_of, len(ids))
            _csr_from_lists([obj[1][k] for k in keys], col_of, len(ids))
        else:
            _csr_from_lists(list(obj.values()), col_of, len(ids))
    t = time.perf_counter() - t0
    out["python_ingest"] = {"seconds": t, "ratings_per_s": total / t, "what": "json.load + per-rating dict lookup (data_reader.py:85-92,134-136)"}
    out["ingest_speedup"] = out["python_ingest"]["seconds"] / out["native_ingest"]["seconds"]
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=3) as pool:
        list(pool.map(lambda a: ingest.load_ratings(d + "native/" + a[0] + ".json", vocab, a[1]), files))
    t = time.perf_counter() - t0
    out["native_ingest_3_threads"] = {"seconds": t, "ratings_per_s": total / t}
    np.random.seed(1)
    t0 = time.perf_counter()
    ls = splitter.split_in_memory(d + "ratings.csv", "movielens")
    t = time.perf_counter() - t0
    out["native_split_in_memory"] = {"seconds": t, "ratings_per_s": n / t,
                                     "what": "CSV -> the reader's train / valid / test stores, no files (%d + %d + %d + %d + %d ratings)"
                                     % (ls.train[1].nnz, ls.valid[1].nnz, ls.valid[3].nnz, ls.test[1].nnz, ls.test[3].nnz)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
