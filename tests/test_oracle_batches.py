"""The oracle's batch restatement vs. the reference's own outputs (tests/golden/batches.npz,
made by tests/golden/make_golden.py from /root/reference/data_reader.py). Bit-exact."""
import numpy as np
import pytest

from oracle import ref_batches
from tests.helpers import golden_batch, oracle_data


def _run_case(case, ds, npz, vectorised):
    data = oracle_data(ds, case["eval_mode"])
    np.random.seed(case["seed"])
    if case["eval_mode"] == "ablation":
        ref_batches.split_rows(data, case["val_split"], seed=case["split_seed"])
        assert np.array_equal(data.train_set, npz[case["id"] + "/train_set"])
        assert np.array_equal(data.val_set, npz[case["id"] + "/val_set"])
        assert np.array_equal(data.test_set, npz[case["id"] + "/test_set"])
    gen = ref_batches.batch_stream(
        data, case["B"], case["sparsity"], train_val_test=case["which"], shuffle=case["shuffle"],
        auxilliary_mask_type=case["aux"], aux_var_value=case["aux_value"],
        return_target_count=case["rtc"], pass_through_input_training=case["pass_through"],
        vectorised=vectorised)
    for n in range(case["n_batches"]):
        item = next(gen)
        feed, targets, tc = golden_batch(npz, case, n)
        assert len(item[0]) == len(feed)
        for got, want in zip(item[0], feed):
            assert got.dtype == np.float64 and np.array_equal(got, want)
        assert np.array_equal(item[1], targets)
        if tc is not None:
            assert item[2] == int(tc)
    assert next(gen) is None and next(gen) is None
    # same number of draws consumed from the global stream as the reference
    assert np.random.random_sample() == float(npz[case["id"] + "/rng_after"])


def test_golden_has_cases(golden_cases):
    assert len(golden_cases) >= 30
    assert all(c["n_batches"] > 0 for c in golden_cases)


@pytest.mark.parametrize("vectorised", [False, True])
def test_oracle_matches_reference(golden_cases, golden_datasets, golden_batches, vectorised):
    for case in golden_cases:
        _run_case(case, golden_datasets[case["dataset"]], golden_batches, vectorised)


def test_draw_split_flags_is_the_choice_stream():
    lengths = [0, 5, 17, 1, 0, 33]
    np.random.seed(9)
    s = np.random.uniform(0.1, 0.8, size=len(lengths))
    want = np.concatenate([np.random.choice([0, 1], size=n, p=[1 - si, si])
                           for n, si in zip(lengths, s)])
    np.random.seed(9)
    got_s, got = ref_batches.draw_split_flags(lengths, [0.1, 0.8])
    assert np.array_equal(got_s, s) and np.array_equal(got, want)
