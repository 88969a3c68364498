"""Host side of the MT19937 jump-ahead (`csrc/ocf_mtjump.h`, no GPU): the characteristic polynomial recovered by
Berlekamp-Massey, x^J mod phi, and the window correlation the device kernel performs, against NumPy's own generator
advanced word by word. NumPy's stream is the reference's stream (data_reader.py:120,130)."""
import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import _lib


def _advance_numpy(key, n_words):
    rs = np.random.RandomState()
    rs.set_state(("MT19937", key.copy(), 624, 0, 0.0))
    rs.bytes(4 * n_words)                       # one 32-bit word per 4 bytes
    st = rs.get_state()
    assert st[2] == 624
    return st[1]


@pytest.mark.parametrize("regens", [1, 2, 31, 32, 33, 256, 7 * 256, 1000, 4096 * 7])
def test_jump_equals_numpy_advanced_by_whole_regenerations(regens):
    lib = _lib.lib()
    rs = np.random.RandomState(20240 + regens)
    rs.random_sample(777)
    key = np.ascontiguousarray(rs.get_state()[1], dtype=np.uint32)
    poly = np.zeros(624, dtype=np.uint32)
    _lib.check(lib.ocf_mt_jump_poly(624 * regens, _lib.ptr(poly)))
    out = np.zeros(624, dtype=np.uint32)
    _lib.check(lib.ocf_mt_jump_apply_host(_lib.ptr(key), _lib.ptr(poly), _lib.ptr(out)))
    want = _advance_numpy(key, 624 * regens)
    assert np.array_equal(out[1:], want[1:])
    assert (out[0] >> 31) == (want[0] >> 31)    # the only bit of word 0 the generator reads
    # and the stream that follows is NumPy's
    a, b = np.random.RandomState(), np.random.RandomState()
    a.set_state(("MT19937", out, 624, 0, 0.0))
    b.set_state(("MT19937", want, 624, 0, 0.0))
    assert np.array_equal(a.random_sample(2000), b.random_sample(2000))


def test_polynomials_compose():
    """x^a * x^b = x^(a+b) mod phi: jumping twice equals jumping once by the sum."""
    lib = _lib.lib()
    key = np.ascontiguousarray(np.random.RandomState(5).get_state()[1], dtype=np.uint32)

    def jump(k, words):
        poly, out = np.zeros(624, dtype=np.uint32), np.zeros(624, dtype=np.uint32)
        _lib.check(lib.ocf_mt_jump_poly(words, _lib.ptr(poly)))
        _lib.check(lib.ocf_mt_jump_apply_host(_lib.ptr(k), _lib.ptr(poly), _lib.ptr(out)))
        return out

    twice = jump(jump(key, 624 * 300), 624 * 500)
    once = jump(key, 624 * 800)
    assert np.array_equal(twice[1:], once[1:]) and (twice[0] >> 31) == (once[0] >> 31)
