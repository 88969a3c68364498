"""GPU parity of the device-resident NumPy stream (block generator workers + polynomial jump-ahead + `k_gather_split<RNG>`):
the keep flags a generator's batches get when MT19937 is replayed on the GPU are bit-identical to
the ones `np.random.uniform` / `np.random.choice` give on the host (data_reader.py:120,130), and
`np.random` ends in the same state."""
import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import synthetic
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader, sync_host_rng

pytestmark = pytest.mark.gpu


def _flags_of_epochs(fs, on_device, B, sparsity, pass_through, seed, epochs=2, shard=None, skip_every=0):
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs, rng_on_device=on_device, shard=shard)
    np.random.seed(seed)
    out = []
    for _ in range(epochs):
        gen = rd.data_gen(B, sparsity, "train", True, None, -1, pass_through_input_training=pass_through)
        k = 0
        while True:
            b = next(gen)
            if b is None:
                break
            k += 1
            if skip_every and k % skip_every == 0:
                continue                      # drawn, never uploaded: the stream must still advance
            out.append((b.rows.copy(), b.flags.copy()))
    sync_host_rng()
    tail = np.random.random_sample(4)          # the host stream continues where the batches left it
    rd.close()
    return out, tail


@pytest.mark.parametrize("B,sparsity,pt", [(8, [0.2, 0.9], False), (61, [0.0, 1.0], True), (128, [0.5, 0.5], False),
                                           (37, [1.0, 1.0], True)])
def test_device_stream_matches_numpy(B, sparsity, pt):
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=4)
    dev, dev_tail = _flags_of_epochs(fs, True, B, sparsity, pt, seed=123)
    host, host_tail = _flags_of_epochs(fs, False, B, sparsity, pt, seed=123)
    assert len(dev) == len(host) and len(dev) > 0
    for (r0, f0), (r1, f1) in zip(dev, host):
        assert np.array_equal(r0, r1)
        assert np.array_equal(f0, f1)
    assert np.array_equal(dev_tail, host_tail)


def test_skipped_batches_advance_the_stream():
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=False, seed=9)
    dev, dev_tail = _flags_of_epochs(fs, True, 32, [0.1, 0.8], False, seed=5, skip_every=3)
    host, host_tail = _flags_of_epochs(fs, False, 32, [0.1, 0.8], False, seed=5, skip_every=3)
    for (r0, f0), (r1, f1) in zip(dev, host):
        assert np.array_equal(r0, r1) and np.array_equal(f0, f1)
    assert np.array_equal(dev_tail, host_tail)


def test_column_shard_reads_its_own_draws():
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=2)
    for shard in [(0, 3), (2, 3)]:
        dev, dev_tail = _flags_of_epochs(fs, True, 50, [0.3, 0.7], False, seed=77, epochs=1, shard=shard)
        host, host_tail = _flags_of_epochs(fs, False, 50, [0.3, 0.7], False, seed=77, epochs=1, shard=shard)
        for (r0, f0), (r1, f1) in zip(dev, host):
            assert np.array_equal(r0, r1) and np.array_equal(f0, f1)
        assert np.array_equal(dev_tail, host_tail)


def test_long_stream_many_regenerations():
    """ML-1M-sized rows: ~10^5 draws per batch cross the 624-word state hundreds of times."""
    fs = synthetic.make_fixed_split("ml1m", reverse_user_item_data=True, seed=0)
    dev, dev_tail = _flags_of_epochs(fs, True, 128, [0.0, 1.0], False, seed=31, epochs=1)
    host, host_tail = _flags_of_epochs(fs, False, 128, [0.0, 1.0], False, seed=31, epochs=1)
    assert sum(f.size for _, f in dev) > 500000
    for (r0, f0), (r1, f1) in zip(dev, host):
        assert np.array_equal(r0, r1) and np.array_equal(f0, f1)
    assert np.array_equal(dev_tail, host_tail)


def test_reseeding_on_the_host_wins():
    """np.random.seed() while the device holds the stream: the next generator starts from the seed."""
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=4)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs, rng_on_device=True)
    np.random.seed(1)
    g = rd.data_gen(16, [0.2, 0.8], "train", True, None, -1)
    next(g).flags
    np.random.seed(99)
    g2 = rd.data_gen(16, [0.2, 0.8], "train", True, None, -1)
    b = next(g2)
    got_rows, got_flags = b.rows.copy(), b.flags.copy()
    rd.close()
    rh = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs, rng_on_device=False)
    np.random.seed(99)
    bh = next(rh.data_gen(16, [0.2, 0.8], "train", True, None, -1))
    assert np.array_equal(got_rows, bh.rows) and np.array_equal(got_flags, bh.flags)
    rh.close()


@pytest.mark.parametrize("workers,block_regens", [(1, 1), (2, 1), (3, 2), (4, 5), (8, 3), (8, 256), (5, 64)])
def test_parallel_generator_workers_give_the_sequential_stream(workers, block_regens):
    """The stream is cut into blocks of `block_regens` regenerations made by `workers` CTAs that jump over each other's
    blocks (GF(2) polynomial jump-ahead). Tiny blocks and a small ring make every batch cross many blocks, every worker
    jump many times and the ring wrap many times; the flags and the final np.random state must not notice."""
    from omnidirectional_collaborative_filtering_b200.data_reader import DeviceRng
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=6)
    sync_host_rng()
    DeviceRng.get().configure(workers, block_regens, 1)
    try:
        dev, dev_tail = _flags_of_epochs(fs, True, 64, [0.1, 0.9], False, seed=workers * 100 + block_regens, skip_every=4)
        info = DeviceRng.get().info()
        assert info["workers"] == workers and info["block_regens"] == block_regens
        host, host_tail = _flags_of_epochs(fs, False, 64, [0.1, 0.9], False, seed=workers * 100 + block_regens, skip_every=4)
        assert len(dev) == len(host) and len(dev) > 4
        for (r0, f0), (r1, f1) in zip(dev, host):
            assert np.array_equal(r0, r1) and np.array_equal(f0, f1)
        assert np.array_equal(dev_tail, host_tail)
        if block_regens <= 5:
            assert info["blocks_enqueued"] > info["ring_blocks"]              # the ring wrapped
    finally:
        sync_host_rng()
        DeviceRng.get().configure(2, 256, 0, pin=False)


def test_ring_grows_for_a_batch_larger_than_the_ring():
    from omnidirectional_collaborative_filtering_b200.data_reader import DeviceRng
    fs = synthetic.make_fixed_split("ml1m", reverse_user_item_data=True, seed=0)
    sync_host_rng()
    DeviceRng.get().configure(4, 2, 1)                    # ring of a few thousand words; a batch needs ~10^5
    try:
        dev, dev_tail = _flags_of_epochs(fs, True, 128, [0.3, 0.6], False, seed=3, epochs=1)
        host, host_tail = _flags_of_epochs(fs, False, 128, [0.3, 0.6], False, seed=3, epochs=1)
        for (r0, f0), (r1, f1) in zip(dev, host):
            assert np.array_equal(r0, r1) and np.array_equal(f0, f1)
        assert np.array_equal(dev_tail, host_tail)
    finally:
        sync_host_rng()
        DeviceRng.get().configure(2, 256, 0, pin=False)
