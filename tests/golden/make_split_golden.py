#!/usr/bin/env python
"""Generate the splitter / file-ingest golden fixtures from the REFERENCE's own scripts.

Run in the build container (needs /root/reference and pandas; neither exists on the GPU box):

    python tests/golden/make_split_golden.py

1. Small ratings CSVs (one per schema case) are written under tests/golden/split/<case>/ratings.csv.
2. `/root/reference/TrainValidTestSplit.py` is executed unmodified except for its module-level parameter
   lines (`full_data_filepath = ...` etc., TrainValidTestSplit.py:17-25), which are re-assigned to point at
   the case; `np.random.seed(seed)` precedes the run. Its output files are stored next to the CSV.
   Cases the script cannot finish under Python 3 (json.dump of np.int64: all-integer CSVs, string ids with
   timestamps) are recorded as such in cases.json with the files it did finish.
3. For the pipeline case the reference's `data_reader.py` (imported as in make_golden.py) reads those files
   from disk and its batches are stored in pipeline_batches.npz; the unique-id lists it needs are written
   here the way TrainValidTestSplit.py:105-118 intends (the script itself crashes there under Python 3).
"""
import contextlib
import io
import json
import os
import re
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import load_reference  # noqa: E402

REF_SPLIT = "/root/reference/TrainValidTestSplit.py"
OUT = os.path.join(HERE, "split")

CASES = [
    # name, schema_type, include_timestamps, reverse_user_item_data, seed
    ("ml", "movielens", False, False, 7),
    ("ml_rev", "movielens", False, True, 8),
    ("ml_ts", "movielens", True, False, 9),
    ("amazon", "amazon", False, False, 10),
    ("netflix_int", "netflix", False, False, 11),        # all-integer rows: the script dies in json.dump
    ("amazon_ts_rev", "amazon", True, True, 12),          # string ids + integer timestamps: same
]


def run_reference_split(params, seed):
    with open(REF_SPLIT) as f:
        src = f.read()
    for k, v in params.items():
        src, n = re.subn(r"(?m)^%s = .*$" % k, "%s = %r" % (k, v), src, count=1)
        assert n == 1, k
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, "TrainValidTestSplit.py", "exec"), {"__name__": "reference_split"})


def make_csv(path, schema, seed, n=320, exotic=True):
    rs = np.random.RandomState(seed)
    users = rs.randint(1, 40, n)
    items = rs.randint(100, 170, n)
    pairs = sorted(set(zip(users.tolist(), items.tolist())))
    rs.shuffle(pairs)
    alphabet = ["1.0", "2.5", "3", "4.5", "5.0", "0.5"]
    if exotic:            # values float32 cannot hold / repr prints in scientific notation (formatting cases only:
        alphabet += ["0.00001", "12345678.125", "1e22", "0.1"]      # the reader stores ratings as float32)
    with open(path, "w", encoding="utf-8") as f:
        f.write("user,movie,rating\n" if schema == "netflix" else "userId,movieId,rating,timestamp\n")
        for u, i in pairs:
            r = alphabet[rs.randint(len(alphabet))]
            ts = 978300000 + rs.randint(0, 10 ** 6)
            if schema == "amazon":
                uid = '"A%d,x"' % u if u % 7 == 0 else "A%dXZ" % u          # a quoted field with a comma
                iid = '"Bé""q%d"' % i if i % 11 == 0 else "B00%d" % i   # non-ASCII + an escaped quote
                f.write("%s,%s,%s,%d\n" % (uid, iid, r, ts))
            elif schema == "netflix":
                f.write("%d,%d,%d\n" % (u, i, 1 + rs.randint(5)))
            else:
                f.write("%d,%d,%s,%d\n" % (u, i, r, ts))


def main():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    record = []
    for name, schema, ts, rev, seed in CASES:
        d = os.path.join(OUT, name) + "/"
        os.makedirs(d + ("reverse_item-user/" if rev else ""))
        make_csv(d + "ratings.csv", schema, seed, exotic=name != "ml")     # "ml" feeds the reader pipeline
        status = "ok"
        try:
            run_reference_split(dict(full_data_filepath=d + "ratings.csv", output_filepath=d, schema_type=schema,
                                     include_timestamps=ts, save_users_and_items=False, reverse_user_item_data=rev), seed)
        except TypeError as e:                    # Object of type int64 is not JSON serializable
            status = "reference failed: %s" % e
        sub = d + ("reverse_item-user/" if rev else "")
        complete = []
        for f in sorted(os.listdir(sub)):
            if f == "ratings.csv" or os.path.isdir(sub + f):
                continue
            if f.endswith(".json"):
                try:
                    with open(sub + f) as fh:
                        json.load(fh)
                except ValueError:
                    os.remove(sub + f)            # the file the script died in
                    continue
            complete.append(f)
        record.append(dict(name=name, schema_type=schema, include_timestamps=ts, reverse_user_item_data=rev, seed=seed,
                           reference=status, files=complete))
        print(name, status, complete)
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(record, f, indent=1)

    # pipeline: reference splitter output -> reference reader (from disk) -> batches
    import pandas as pd
    d = os.path.join(OUT, "ml") + "/"
    ratings = pd.read_csv(d + "ratings.csv")
    with open(d + "unique_items_list.json", "w") as f:
        json.dump([int(x) for x in ratings["movieId"].unique()], f)
    with open(d + "unique_users_list.json", "w") as f:
        json.dump([str(int(x)) for x in ratings["userId"].unique()], f)
    ref = load_reference()
    n_items = int(ratings["movieId"].nunique())
    with open(d + "ratingsByUser_dicts_train.json") as f:
        n_rows = len(json.load(f))
    rd = ref.data_reader(n_items, n_rows, d, nonsequentialusers=False, use_json=True, eval_mode="fixed_split",
                         useTimestamps=False, reverse_user_item_data=False)
    rd.train_set, rd.val_set, rd.test_set = list(rd.train_set), list(rd.val_set), list(rd.test_set)   # Py2 lists
    store = {"n_items": np.asarray(n_items), "n_rows": np.asarray(n_rows)}
    for which, sparsity, aux, pt, seed in (("train", [0.3, 0.8], "dropout", False, 41), ("valid", None, None, False, 42),
                                           ("test", None, "both", False, 43)):
        np.random.seed(seed)
        gen = rd.data_gen(8, sparsity, train_val_test=which, shuffle=True, auxilliary_mask_type=aux, aux_var_value=-1,
                          return_target_count=which != "train", pass_through_input_training=pt)
        n = 0
        while True:
            item = next(gen)
            if item is None:
                break
            for k, arr in enumerate(item[0]):
                store["%s/b%d/in%d" % (which, n, k)] = np.asarray(arr, dtype=np.float64)
            store["%s/b%d/targets" % (which, n)] = np.asarray(item[1], dtype=np.float64)
            if len(item) > 2:
                store["%s/b%d/target_count" % (which, n)] = np.asarray(item[2])
            n += 1
        store["%s/n_batches" % which] = np.asarray(n)
        print("pipeline", which, "batches:", n)
    # ablation mode from disk: `ratingsByUser_dict.json` (data_reader.py:55; no script of the reference writes it - it
    # has the layout of the train file) + a `unique_users_list.json` in another order than the dict's keys
    a = os.path.join(OUT, "ml_ablation") + "/"
    os.makedirs(a)
    shutil.copy(d + "ratingsByUser_dicts_train.json", a + "ratingsByUser_dict.json")
    shutil.copy(d + "unique_items_list.json", a + "unique_items_list.json")
    with open(a + "ratingsByUser_dict.json") as f:
        users = sorted(json.load(f).keys(), key=lambda k: (len(k), k))[::-1]
    with open(a + "unique_users_list.json", "w") as f:
        json.dump(users, f)
    rd = ref.data_reader(n_items, len(users), a, nonsequentialusers=True, use_json=True, eval_mode="ablation",
                         useTimestamps=False, reverse_user_item_data=False)
    np.random.seed(51)
    rd.split_for_validation([0.6, 0.2, 0.2])
    store["ablation/n_users"] = np.asarray(len(users))
    for which, sparsity, seed in (("train", [0.2, 0.7], 52), ("test", [0.5, 0.5], 53)):
        np.random.seed(seed)
        gen = rd.data_gen(4, sparsity, train_val_test=which, shuffle=True, auxilliary_mask_type="causal", aux_var_value=-1)
        n = 0
        while True:
            item = next(gen)
            if item is None:
                break
            for k, arr in enumerate(item[0]):
                store["ablation/%s/b%d/in%d" % (which, n, k)] = np.asarray(arr, dtype=np.float64)
            store["ablation/%s/b%d/targets" % (which, n)] = np.asarray(item[1], dtype=np.float64)
            n += 1
        store["ablation/%s/n_batches" % which] = np.asarray(n)
        print("ablation", which, "batches:", n)
    np.savez_compressed(os.path.join(OUT, "pipeline_batches.npz"), **store)


if __name__ == "__main__":
    main()
