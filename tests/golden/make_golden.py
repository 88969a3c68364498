#!/usr/bin/env python
"""Generate the golden batch fixtures from the REFERENCE's own `data_reader.py`.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports `/root/reference/data_reader.py` unmodified. Three accommodations, none of which
touches the arithmetic:
  1. a stub `tensorflow` module exposing `SparseTensor` (imported at data_reader.py:7, never used),
  2. `load_data` (data_reader.py:85-92) overridden to serve in-memory dicts instead of files,
  3. `train_set/val_set/test_set` turned into lists (data_reader.py:78-80 are Python-2 lists;
     Python-3 dict views cannot be permuted/indexed at :327/:227).

Outputs (committed):
  tests/golden/dataset_<name>.json   the rating dicts the cases run on (incl. duplicate
                                     ratings inside a row, None input rows, empty rows)
  tests/golden/cases.json            one entry per case: generator arguments + seed
  tests/golden/batches.npz           every array of every batch of every case, float64,
                                     plus the RNG stream position after the case
"""
from __future__ import print_function

import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from omnidirectional_collaborative_filtering_b200 import synthetic  # noqa: E402

REF = "/root/reference/data_reader.py"


def load_reference():
    tf_stub = types.ModuleType("tensorflow")
    tf_stub.SparseTensor = object
    sys.modules.setdefault("tensorflow", tf_stub)
    spec = importlib.util.spec_from_file_location("ref_data_reader", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_dataset(reverse, seed):
    """Tiny fixed-split dataset in reference dict form, with hand-injected edge cases."""
    shape = synthetic.SHAPES["tiny"]
    u, i, r = synthetic.make_ratings(shape, seed)
    fs = synthetic.build_fixed_split(u, i, r, shape.n_users, shape.n_items, reverse, seed + 1)
    d = synthetic.to_reference_dicts(fs)
    train, (va_in, va_tg), (te_in, te_tg) = d["train"], d["valid"], d["test"]
    cols = d["unique_cols"]
    tkeys = list(train.keys())
    # (a) duplicate ratings inside train rows: same column, different value, both orders
    train[tkeys[0]].append([train[tkeys[0]][0][0], 2.5])
    train[tkeys[1]].insert(0, [train[tkeys[1]][-1][0], 0.5])
    train[tkeys[2]].extend([[train[tkeys[2]][1][0], 4.5], [train[tkeys[2]][1][0], 1.5]])
    # (b) an empty train row
    train[tkeys[3]] = []
    vkeys = list(va_tg.keys())
    # (c) None input rows (row has no train rating, TrainValidTestSplit.py:193)
    va_in[vkeys[0]] = None
    va_in[vkeys[5]] = None
    # (d) a target that is also an input of the same row, and a duplicated target
    if va_in[vkeys[1]]:
        va_tg[vkeys[1]].append([va_in[vkeys[1]][0][0], 3.5])
    va_tg[vkeys[2]].append([va_tg[vkeys[2]][0][0], 1.0])
    # (e) duplicated input
    if va_in[vkeys[3]]:
        va_in[vkeys[3]] = list(va_in[vkeys[3]]) + [[va_in[vkeys[3]][0][0], 0.25]]
    # (f) a rating of exactly 0.0 as input and as target (Jester has them)
    train[tkeys[4]][0][1] = 0.0
    va_tg[vkeys[4]][0][1] = 0.0
    tekeys = list(te_tg.keys())
    te_in[tekeys[0]] = None
    n_rows = shape.n_items if reverse else shape.n_users
    n_cols = shape.n_users if reverse else shape.n_items
    # ablation-mode dict: every row, keyed by the raw (string) row id; rows are the union of
    # the train rows plus a few rows that only exist in valid/test (given train-like lists)
    abl = {}
    for k in tkeys:
        abl[k] = train[k]
    for k in list(va_tg.keys()) + list(te_tg.keys()):
        if k not in abl:
            abl[k] = va_tg.get(k) or te_tg.get(k)
    return {
        "reverse": bool(reverse), "n_rows": n_rows, "n_cols": n_cols,
        "unique_cols": cols, "unique_rows": list(abl.keys()),
        "train": train, "valid": [va_in, va_tg], "test": [te_in, te_tg], "ablation": abl,
    }


def reference_reader(ref, ds, eval_mode):
    files = {
        "unique_items_list": ds["unique_cols"], "unique_users_list": ds["unique_rows"],
        "ratingsByUser_dict": ds["ablation"],
        "ratingsByUser_dicts_train": ds["train"],
        "ratingsByUser_dicts_valid": ds["valid"],
        "ratingsByUser_dicts_test": ds["test"],
    }

    class Reader(ref.data_reader):
        def load_data(self, filepath, filename, use_json):
            return files[filename]

    n_rows = len(ds["unique_rows"]) if eval_mode == "ablation" else ds["n_rows"]
    rd = Reader(ds["n_cols"], n_rows, "", nonsequentialusers=True, use_json=True,
                eval_mode=eval_mode, useTimestamps=False, reverse_user_item_data=False)
    if eval_mode == "fixed_split":
        rd.train_set, rd.val_set, rd.test_set = list(rd.train_set), list(rd.val_set), list(rd.test_set)
    return rd


def cases():
    out = []
    cid = 0

    def add(**kw):
        nonlocal cid
        kw["id"] = "c%03d" % cid
        cid += 1
        out.append(kw)

    # fixed_split / train: aux types x pass-through x sparsity ranges
    for aux in ["dropout", "causal", "zeros", "both", None]:
        for pt in (False, True):
            add(dataset="rev", eval_mode="fixed_split", which="train", B=8, sparsity=[0.2, 0.9],
                shuffle=True, aux=aux, aux_value=-1, pass_through=pt, seed=11, rtc=False)
    for sp in ([1.0, 1.0], [0.0, 0.0], [0.5, 0.5], [0.0, 1.0]):
        add(dataset="rev", eval_mode="fixed_split", which="train", B=8, sparsity=sp,
            shuffle=True, aux="dropout", aux_value=-1, pass_through=True, seed=5, rtc=False)
        add(dataset="fwd", eval_mode="fixed_split", which="train", B=16, sparsity=sp,
            shuffle=False, aux=None, aux_value=1, pass_through=False, seed=6, rtc=False)
    # fixed_split / valid + test
    for aux in ["dropout", "causal", "zeros", "both", None]:
        add(dataset="rev", eval_mode="fixed_split", which="valid", B=8, sparsity=[1.0, 1.0],
            shuffle=True, aux=aux, aux_value=-1, pass_through=False, seed=21, rtc=True)
    add(dataset="fwd", eval_mode="fixed_split", which="valid", B=4, sparsity=None,
        shuffle=False, aux="both", aux_value=2.5, pass_through=False, seed=22, rtc=False)
    add(dataset="rev", eval_mode="fixed_split", which="test", B=8, sparsity=None,
        shuffle=True, aux=None, aux_value=-1, pass_through=False, seed=23, rtc=True)
    add(dataset="fwd", eval_mode="fixed_split", which="test", B=16, sparsity=None,
        shuffle=True, aux="causal", aux_value=-1, pass_through=False, seed=24, rtc=True)
    # ablation: row split then train / valid / test batches all go through the random split
    for which, sp in (("train", [0.5, 0.5]), ("valid", [0.5, 0.5]), ("test", [0.0, 0.0]),
                      ("test", [0.1, 0.1]), ("test", [0.9, 0.9])):
        add(dataset="rev", eval_mode="ablation", which=which, B=4, sparsity=sp, shuffle=True,
            aux="dropout", aux_value=-1, pass_through=False, seed=31, rtc=False,
            val_split=[0.5, 0.25, 0.25], split_seed=3)
    add(dataset="fwd", eval_mode="ablation", which="train", B=8, sparsity=[0.0, 1.0], shuffle=False,
        aux="both", aux_value=-1, pass_through=True, seed=32, rtc=False,
        val_split=[0.6, 0.2, 0.2], split_seed=None)
    return out


def main():
    ref = load_reference()
    datasets = {"rev": make_dataset(True, 100), "fwd": make_dataset(False, 200)}
    for name, ds in datasets.items():
        with open(os.path.join(HERE, "dataset_%s.json" % name), "w") as f:
            json.dump(ds, f)
    store = {}
    all_cases = cases()
    for c in all_cases:
        ds = datasets[c["dataset"]]
        rd = reference_reader(ref, ds, c["eval_mode"])
        np.random.seed(c["seed"])
        if c["eval_mode"] == "ablation":
            rd.split_for_validation(c["val_split"], seed=c["split_seed"])
            store[c["id"] + "/train_set"] = np.asarray(rd.train_set)
            store[c["id"] + "/val_set"] = np.asarray(rd.val_set)
            store[c["id"] + "/test_set"] = np.asarray(rd.test_set)
        gen = rd.data_gen(c["B"], c["sparsity"], train_val_test=c["which"], shuffle=c["shuffle"],
                          auxilliary_mask_type=c["aux"], aux_var_value=c["aux_value"],
                          return_target_count=c["rtc"],
                          pass_through_input_training=c["pass_through"])
        n = 0
        while True:
            item = next(gen)
            if item is None:
                break
            for k, arr in enumerate(item[0]):
                store["%s/b%d/in%d" % (c["id"], n, k)] = np.asarray(arr, dtype=np.float64)
            store["%s/b%d/targets" % (c["id"], n)] = np.asarray(item[1], dtype=np.float64)
            if len(item) > 2:
                store["%s/b%d/target_count" % (c["id"], n)] = np.asarray(item[2])
            n += 1
        assert next(gen) is None                  # stays None (data_reader.py:418-419)
        c["n_batches"] = n
        store[c["id"] + "/rng_after"] = np.asarray(np.random.random_sample())
        print(c["id"], c["eval_mode"], c["which"], "batches:", n)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(all_cases, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "batches.npz"), **store)
    print("wrote", len(store), "arrays")


if __name__ == "__main__":
    main()
