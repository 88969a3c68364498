"""Host logic of the product's data_reader mirror (no GPU): set handling, RNG replay, batch
order and flags, checked against the reference's golden batches by densifying the (row ids,
flags) plan on the host."""
import numpy as np

from tests.helpers import golden_batch, host_densify, product_reader


def test_reader_plans_match_reference(golden_cases, golden_datasets, golden_batches):
    for case in golden_cases:
        ds = golden_datasets[case["dataset"]]
        rd = product_reader(ds, case["eval_mode"])
        np.random.seed(case["seed"])
        if case["eval_mode"] == "ablation":
            rd.split_for_validation(case["val_split"], seed=case["split_seed"])
            assert np.array_equal(rd.train_set, golden_batches[case["id"] + "/train_set"])
            assert np.array_equal(rd.test_set, golden_batches[case["id"] + "/test_set"])
        else:
            assert rd.train_set == list(ds["train"].keys())
            assert (rd.train_set_size, rd.val_set_size, rd.test_set_size) == \
                (len(ds["train"]), len(ds["valid"][1]), len(ds["test"][1]))
        gen = rd.data_gen(case["B"], case["sparsity"], train_val_test=case["which"], shuffle=case["shuffle"],
                          auxilliary_mask_type=case["aux"], aux_var_value=case["aux_value"],
                          return_target_count=case["rtc"], pass_through_input_training=case["pass_through"])
        for n in range(case["n_batches"]):
            batch = next(gen)
            feed, targets, tc = golden_batch(golden_batches, case, n)
            got_feed, got_t = host_densify(batch)
            assert len(got_feed) == len(feed), case["id"]
            for g, w in zip(got_feed, feed):
                assert np.array_equal(g, w), (case["id"], n)
            assert np.array_equal(got_t, targets)
            assert len(batch) == (3 if tc is not None else 2)
            if tc is not None:
                assert batch.target_count == int(tc)
        assert next(gen) is None and next(gen) is None
        assert np.random.random_sample() == float(golden_batches[case["id"] + "/rng_after"]), case["id"]


def test_scalar_sparsity_is_accepted(golden_datasets):
    rd = product_reader(golden_datasets["rev"], "ablation")
    rd.split_for_validation([0.5, 0.25, 0.25], seed=1)
    np.random.seed(4)
    a = next(rd.data_gen(4, 0.4, "test"))
    np.random.seed(4)
    b = next(rd.data_gen(4, [0.4, 0.4], "test"))
    assert np.array_equal(a.flags, b.flags) and np.array_equal(a.rows, b.rows)


def test_max_batch_entries_bounds_every_batch(golden_datasets):
    rd = product_reader(golden_datasets["fwd"], "fixed_split")
    cap = rd.max_batch_entries(8)
    for which in ("train", "valid", "test"):
        gen = rd.data_gen(8, [0.5, 0.5], which)
        while True:
            b = next(gen)
            if b is None:
                break
            assert b.n_entries <= cap
