"""Shared helpers for the test-suite (oracle-side construction of reference-shaped data)."""
import numpy as np

from oracle import ref_batches


def oracle_data(ds, eval_mode):
    """RefData over one of the golden datasets, wired like tests/golden/make_golden.py wires
    the reference reader."""
    if eval_mode == "ablation":
        return ref_batches.RefData(ds["n_cols"], len(ds["unique_rows"]), ds["unique_cols"],
                                   eval_mode="ablation", user_dict=ds["ablation"],
                                   nonsequentialusers=True, unique_rows=ds["unique_rows"])
    return ref_batches.RefData(ds["n_cols"], ds["n_rows"], ds["unique_cols"],
                               eval_mode="fixed_split", train=ds["train"],
                               valid=tuple(ds["valid"]), test=tuple(ds["test"]),
                               nonsequentialusers=True, unique_rows=ds["unique_rows"])


def golden_batch(npz, case, n):
    """(input list, targets, target_count or None) of batch n of a golden case."""
    cid = case["id"]
    feed = []
    k = 0
    while "%s/b%d/in%d" % (cid, n, k) in npz.files:
        feed.append(npz["%s/b%d/in%d" % (cid, n, k)])
        k += 1
    tc = npz["%s/b%d/target_count" % (cid, n)] if "%s/b%d/target_count" % (cid, n) in npz.files else None
    return feed, npz["%s/b%d/targets" % (cid, n)], tc
