"""Host bookkeeping of `data_reader.DeviceRng` (no GPU): tickets, in-order consumption with
skipped batches, hand-back of the stream to `np.random`, and a host reseed winning over the device
copy. The C library is replaced by a stand-in whose `ocf_rng_*` calls advance a NumPy RandomState,
which is exactly what the device kernel does (tests/test_gpu_rng.py proves that part on the GPU)."""
import ctypes as C

import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import _lib, data_reader as dr


class _FakeLib(object):
    def __init__(self):
        self.rs = np.random.RandomState(0)
        self.skips = []

    def ocf_device_count(self):
        return 1

    def ocf_rng_create(self, out):
        out._obj.value = 1234
        return 0

    def ocf_rng_set_state(self, handle, key_ptr, pos):
        key = np.ctypeslib.as_array((C.c_uint32 * 624).from_address(key_ptr.value)).copy()
        self.rs.set_state(("MT19937", key, int(pos), 0, 0.0))
        return 0

    def ocf_rng_skip(self, handle, n):
        self.skips.append(int(n))
        self.rs.random_sample(int(n))
        return 0

    def ocf_rng_prefetch(self, handle, n):       # a hint: the real library starts generating ahead
        self.prefetched = getattr(self, "prefetched", []) + [int(n)]
        return 0

    def ocf_rng_configure(self, handle, workers, block_regens, ring_words_min):
        self.layout = (int(workers), int(block_regens))
        return 0

    def ocf_rng_info(self, handle, buf):
        w, c = getattr(self, "layout", (2, 256))
        for k, v in enumerate((w, c, 16, 16 * 624 * c, 0, 0)):
            buf[k] = v
        return 0

    def ocf_rng_get_state(self, handle, key_ptr, pos):
        st = self.rs.get_state()
        np.ctypeslib.as_array((C.c_uint32 * 624).from_address(key_ptr.value))[:] = st[1]
        pos._obj.value = int(st[2])
        return 0


@pytest.fixture
def fake(monkeypatch):
    lib = _FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: lib)
    monkeypatch.setattr(_lib, "require_gpu", lambda: None)
    monkeypatch.setattr(dr.DeviceRng, "_instance", None)
    yield lib
    dr.DeviceRng._instance = None


def test_tickets_skips_and_release_follow_the_host_stream(fake):
    np.random.seed(42)
    np.random.random_sample(5)                    # the stream is somewhere in the middle
    want = np.random.RandomState()
    want.set_state(np.random.get_state())
    rng = dr.DeviceRng.get()
    draws = [100, 37, 250, 9]
    tickets = [rng.ticket(n) for n in draws]
    # batch 0 is uploaded, batch 1 never is, batch 2 is uploaded: its upload skips batch 1's draws first
    assert rng.consume(tickets[0]) is not None
    fake.rs.random_sample(draws[0])               # what ocf_batch_fill_split_rng does on the device
    rng.consume(tickets[2])
    assert fake.skips == [37]
    fake.rs.random_sample(draws[2])
    with pytest.raises(RuntimeError):
        rng.consume(tickets[1])                   # its draws are gone
    dr.sync_host_rng()                            # batch 3 was drawn but never uploaded: skipped at release
    assert fake.skips == [37, 9]
    want.random_sample(sum(draws))
    assert np.random.random_sample() == want.random_sample()
    assert rng.host_state is None and not rng.pending


def test_reseeding_on_the_host_discards_the_device_copy(fake):
    np.random.seed(1)
    rng = dr.DeviceRng.get()
    t = rng.ticket(10)
    rng.consume(t)
    fake.rs.random_sample(10)
    np.random.seed(7)                             # the caller reseeds while the device holds the stream
    expect = np.random.RandomState(7).random_sample()
    dr.sync_host_rng()
    assert np.random.random_sample() == expect
    assert rng.host_state is None


def test_prefetch_runs_at_most_a_few_batches_ahead(fake, monkeypatch):
    """The generator thread may have drawn tickets for dozens of batches; the device workers are asked to run only
    OCF_RNG_AHEAD batches ahead of the one being uploaded (what is in flight has to drain whenever the stream goes
    back to the host, i.e. at every epoch start)."""
    monkeypatch.delenv("OCF_RNG_AHEAD", raising=False)
    np.random.seed(1)
    rng = dr.DeviceRng.get()
    tickets = [rng.ticket(1000) for _ in range(40)]
    rng.consume(tickets[0])
    assert fake.prefetched[-1] == 3 * 1000                      # 40 batches pending, default cap 3
    fake.rs.random_sample(1000)
    monkeypatch.setenv("OCF_RNG_AHEAD", "5")
    rng.consume(tickets[1])
    assert fake.prefetched[-1] == 5 * 1000
    fake.rs.random_sample(1000)
    for t in tickets[2:39]:
        rng.consume(t)
        fake.rs.random_sample(1000)
    monkeypatch.delenv("OCF_RNG_AHEAD")
    rng.consume(tickets[39])                                    # nothing pending behind it: its own draws only
    assert fake.prefetched[-1] == 1000
