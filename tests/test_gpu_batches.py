"""GPU parity, batch construction (kernel K1 through the C ABI): the dense arrays scattered
from the device tiles must equal the reference's own outputs bit for bit."""
import numpy as np
import pytest

from tests.helpers import golden_batch, product_reader

pytestmark = pytest.mark.gpu


def test_gathered_batches_equal_reference(golden_cases, golden_datasets, golden_batches):
    for case in golden_cases:
        rd = product_reader(golden_datasets[case["dataset"]], case["eval_mode"])
        np.random.seed(case["seed"])
        if case["eval_mode"] == "ablation":
            rd.split_for_validation(case["val_split"], seed=case["split_seed"])
        gen = rd.data_gen(case["B"], case["sparsity"], train_val_test=case["which"], shuffle=case["shuffle"],
                          auxilliary_mask_type=case["aux"], aux_var_value=case["aux_value"],
                          return_target_count=case["rtc"], pass_through_input_training=case["pass_through"])
        for n in range(case["n_batches"]):
            item = next(gen)
            feed, targets, tc = golden_batch(golden_batches, case, n)
            if tc is None:
                got_feed, got_t = item               # unpacks like the reference's tuples
            else:
                got_feed, got_t, got_tc = item
                assert got_tc == int(tc)
            assert len(got_feed) == len(feed)
            for g, w in zip(got_feed, feed):
                assert g.dtype == np.float64 and np.array_equal(g, w), (case["id"], n)
            assert np.array_equal(got_t, targets), (case["id"], n)
        assert next(gen) is None
        rd.close()


def test_store_reports_duplicates(golden_datasets):
    rd = product_reader(golden_datasets["rev"], "fixed_split")
    info = rd.store("train").info()
    assert info["has_dups"] == 1 and info["nnz"] == rd.store("train").csr.nnz
    rd.close()


def test_large_batch_round_trip():
    """Full-size property: densify(x) + densify(t) of a split batch reproduces the rows of the
    store (every rating lands in exactly one of input/target when not pass-through)."""
    from omnidirectional_collaborative_filtering_b200 import synthetic
    from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=3)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    np.random.seed(0)
    gen = rd.data_gen(128, [0.3, 0.7], "train", pass_through_input_training=False, auxilliary_mask_type="causal")
    batch = next(gen)
    (x, observed, mask_out), t = batch
    dense = np.zeros_like(x)
    for b, row in enumerate(batch.rows):
        c, v = fs.train.row(int(row))
        dense[b, c] = v
    assert np.array_equal(x + t, dense)
    assert np.array_equal(observed != 0, dense != 0)
    assert np.array_equal(mask_out != 0, t != 0)
    rd.close()
