"""ORACLE (test infrastructure): Philox4x32-10, the counter-based generator that defines the
hidden-layer dropout masks shared by the oracle and the CUDA kernels.

The reference takes its dropout masks from TensorFlow's RNG (`model.py:72-73`,
`Dropout(p, noise_shape=[B, H])`), which cannot be reproduced outside TF 1.3. Both sides of
the parity tests therefore use this specification instead:

    counter = (unit >> 2, batch_row, layer, step)      key = (seed_lo, seed_hi)
    word    = philox4x32_10(counter, key)[unit & 3]
    keep    = (word >> 8) >= floor(p * 2**24)          # P(drop) = floor(p*2^24) / 2^24
    h_out   = keep ? h / (1 - p) : 0                   # inverted dropout, like Keras

Only `tests/`, `smoke()` and `bench.py`'s CPU legs may import this module.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over counter arrays (uint32 each); key words are Python ints. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def dropout_keep(seed, step, layer, n_rows, n_units, p, row0=0):
    """uint8 [n_rows, n_units] keep-mask (1 = kept) for one layer of one step."""
    units4 = np.arange((n_units + 3) // 4, dtype=np.uint32)[None, :]
    rows = (np.arange(n_rows, dtype=np.uint32) + np.uint32(row0))[:, None]
    words = philox4x32_10(units4, rows, np.uint32(layer), np.uint32(step),
                          int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    w = np.stack(words, axis=-1).reshape(n_rows, -1)[:, :n_units]
    thresh = np.uint32(int(np.floor(float(p) * 16777216.0)))
    return ((w >> np.uint32(8)) >= thresh).astype(np.uint8)
