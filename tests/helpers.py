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


def product_reader(ds, eval_mode):
    """The product's data_reader over a golden dataset (in-memory 'files')."""
    from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
    files = {
        "unique_items_list": ds["unique_cols"], "unique_users_list": ds["unique_rows"],
        "ratingsByUser_dict": ds["ablation"],
        "ratingsByUser_dicts_train": ds["train"],
        "ratingsByUser_dicts_valid": ds["valid"],
        "ratingsByUser_dicts_test": ds["test"],
    }
    n_rows = len(ds["unique_rows"]) if eval_mode == "ablation" else ds["n_rows"]
    return data_reader(ds["n_cols"], n_rows, "", nonsequentialusers=True, use_json=True,
                       eval_mode=eval_mode, useTimestamps=False, reverse_user_item_data=False, data=files)


def host_densify(batch):
    """Dense arrays a product Batch stands for, computed on the host straight from its row ids and
    keep flags (test-side restatement of data_reader.py:158-169 / :234-268; no GPU involved)."""
    B, N, a = batch.n_rows, batch.n_cols, batch.aux_value
    mask_in, mask_out, x, t, observed = (np.zeros([B, N]) for _ in range(5))
    pos = 0
    for b, row in enumerate(batch.rows):
        if batch.kind == "split":
            cols, vals = batch.source.csr.row(int(row))
            f = batch.flags[pos:pos + len(cols)].astype(bool)
            pos += len(cols)
            for c, v, fi in zip(cols, vals, f):
                if fi:
                    mask_in[b, c] = a; x[b, c] = v
                    if batch.pass_through:
                        mask_out[b, c] = a; t[b, c] = v
                else:
                    mask_out[b, c] = a; t[b, c] = v
                observed[b, c] = a
        else:
            cols, vals = batch.source.in_store.csr.row(int(row))
            for c, v in zip(cols, vals):
                mask_in[b, c] = a; x[b, c] = v; observed[b, c] = a
            cols, vals = batch.source.tgt_store.csr.row(int(row))
            for c, v in zip(cols, vals):
                mask_out[b, c] = a; t[b, c] = v; observed[b, c] = a
    return ref_batches.feed_list((mask_in, mask_out, x, t, observed), batch.aux_type), t
