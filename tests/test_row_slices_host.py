"""Row (data) parallelism, host side (no GPU): a rank's `Batch.row_slice` is exactly its rows of the
global batch - same rows, same keep flags, same dense arrays as the reference builds for those rows
(`data_reader.py:95-298`) - for split and fixed-split batches."""
import numpy as np
import pytest

from tests.helpers import golden_batch, host_densify, product_reader


@pytest.mark.parametrize("world", [2, 4])
def test_row_slices_partition_the_reference_batch(golden_cases, golden_datasets, golden_batches, world):
    done = 0
    for case in golden_cases:
        if case["B"] % world or case["eval_mode"] != "fixed_split":
            continue
        ds = golden_datasets[case["dataset"]]
        rd = product_reader(ds, case["eval_mode"])
        rd.rng_on_device = False                       # host stream: this test must run without a GPU
        np.random.seed(case["seed"])
        gen = rd.data_gen(case["B"], case["sparsity"], train_val_test=case["which"], shuffle=case["shuffle"],
                          auxilliary_mask_type=case["aux"], aux_var_value=case["aux_value"],
                          return_target_count=case["rtc"], pass_through_input_training=case["pass_through"])
        per = case["B"] // world
        for n in range(case["n_batches"]):
            batch = next(gen)
            feed, targets, tc = golden_batch(golden_batches, case, n)
            count = 0
            for rank in range(world):
                sub = batch.row_slice(rank, world)
                assert sub.row0 == rank * per and sub.rows_total == case["B"] and sub.n_rows == per
                assert np.array_equal(sub.rows, batch.rows[rank * per:(rank + 1) * per])
                got_feed, got_t = host_densify(sub)
                for g, w in zip(got_feed, feed):
                    assert np.array_equal(g, w[rank * per:(rank + 1) * per]), (case["id"], rank, n)
                assert np.array_equal(got_t, targets[rank * per:(rank + 1) * per])
                count += sub.n_ratings
            assert count == batch.n_ratings
        done += 1
    assert done >= 3


def test_row_slice_rejects_ragged_split(golden_datasets):
    ds = golden_datasets["rev"]
    rd = product_reader(ds, "fixed_split")
    rd.rng_on_device = False
    np.random.seed(0)
    batch = next(rd.data_gen(6, [0.5, 0.5], "train", True, None, -1))
    with pytest.raises(ValueError):
        batch.row_slice(0, 4)
