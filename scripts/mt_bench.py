#!/usr/bin/env python
"""Rate of the device-resident MT19937 stream (k_mt_words, one CTA): draws/s over a long skip,
wall clock around `ocf_rng_skip` + the synchronising `ocf_rng_get_state`, checked against NumPy."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnidirectional_collaborative_filtering_b200 import _lib

lib = _lib.lib()
rng = C.c_void_p()
_lib.check(lib.ocf_rng_create(C.byref(rng)))
rs = np.random.RandomState(7)
st = rs.get_state()
key = np.ascontiguousarray(st[1], dtype=np.uint32)
_lib.check(lib.ocf_rng_set_state(rng, _lib.ptr(key), int(st[2])))
out_key, pos = np.empty(624, dtype=np.uint32), C.c_int32()
_lib.check(lib.ocf_rng_skip(rng, 1000))
_lib.check(lib.ocf_rng_get_state(rng, _lib.ptr(out_key), C.byref(pos)))          # warm-up + sync
n = 20_000_000
t0 = time.perf_counter()
_lib.check(lib.ocf_rng_skip(rng, n))
_lib.check(lib.ocf_rng_get_state(rng, _lib.ptr(out_key), C.byref(pos)))
dt = time.perf_counter() - t0
cyc, ns = C.c_int64(), C.c_int64()
_lib.check(lib.ocf_rng_last_timing(rng, C.byref(cyc), C.byref(ns)))
print("in-kernel: %d SM cycles, %.2f ms -> %.0f MHz, %.0f cycles per regeneration" %
      (cyc.value, ns.value / 1e6, cyc.value / max(ns.value, 1) * 1e3, cyc.value / (2 * n / 624)))
rs.random_sample(1000 + n)
want = rs.get_state()
ok = np.array_equal(want[1], out_key) and want[2] == pos.value
print("k_mt_words: %d draws in %.2f ms -> %.2f G draws/s, %.3f us per 624-word regeneration; state matches NumPy: %s"
      % (n, dt * 1e3, n / dt / 1e9, dt * 1e6 / (2 * n / 624), ok))
sys.exit(0 if ok else 1)
