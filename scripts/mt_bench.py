"""Throughput of the device-resident NumPy MT19937 stream (block generator workers + polynomial jump-ahead):
wall clock around `ocf_rng_prefetch` of n draws + the synchronising `ocf_rng_get_state` after a skip to their end,
for 1, 2, 4 and 8 worker CTAs, checked against NumPy.   python scripts/mt_bench.py [n_draws]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnidirectional_collaborative_filtering_b200 import _lib

lib = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
ref = np.random.RandomState(7)
st = ref.get_state()
key = np.ascontiguousarray(st[1], dtype=np.uint32)
ref.random_sample(1000)
ref.random_sample(n)
want = ref.get_state()
for workers in (1, 2, 4, 8):
    rng = C.c_void_p()
    _lib.check(lib.ocf_rng_create(C.byref(rng)))
    _lib.check(lib.ocf_rng_configure(rng, workers, 256, 2 * n + 4096))
    _lib.check(lib.ocf_rng_set_state(rng, _lib.ptr(key), int(st[2])))
    out_key, pos = np.empty(624, dtype=np.uint32), C.c_int32()
    _lib.check(lib.ocf_rng_prefetch(rng, 1000))
    _lib.check(lib.ocf_rng_skip(rng, 1000))
    _lib.check(lib.ocf_rng_get_state(rng, _lib.ptr(out_key), C.byref(pos)))          # warm-up + sync
    t0 = time.perf_counter()
    _lib.check(lib.ocf_rng_prefetch(rng, n))
    _lib.check(lib.ocf_rng_skip(rng, n))
    _lib.check(lib.ocf_rng_get_state(rng, _lib.ptr(out_key), C.byref(pos)))
    dt = time.perf_counter() - t0
    cyc, ns = C.c_int64(), C.c_int64()
    _lib.check(lib.ocf_rng_last_timing(rng, C.byref(cyc), C.byref(ns)))
    ok = np.array_equal(out_key, want[1]) and pos.value == want[2]
    print("workers %d: %d draws in %.3f ms = %.2f G draws/s (block kernel: %d cycles, %.1f us per 256 regenerations) state %s"
          % (workers, n, dt * 1e3, n / dt / 1e9, cyc.value, ns.value / 1e3, "== NumPy" if ok else "MISMATCH"))
    lib.ocf_rng_destroy(rng)
