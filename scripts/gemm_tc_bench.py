"""Where the time of one k_gemm_tc launch goes (csrc/ocf_gemm_tc.cuh): CUDA-event time per launch over back-to-back
launches and the %globaltimer phase stamps of one CTA, for the shapes the training steps use.
    python scripts/gemm_tc_bench.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnidirectional_collaborative_filtering_b200 import _lib

lib = _lib.lib()
names = ["entry", "setup", "operands", "products", "parked+sync", "epilogue", "sync2", "exit"]
shapes = [("ml20m forward   W[1024,512] h[128,1024]", 1, 0, 512, 128, 1024, 3),
          ("ml20m backward  W[1024,512] dz[128,512]", 0, 0, 1024, 128, 512, 1),
          ("ml20m gradient  dz[128,512] h[128,1024]", 1, 1, 512, 1024, 128, 2),
          ("jester forward  W[256,256] h[128,256]", 1, 0, 256, 128, 256, 3),
          ("jester gradient dz[128,256] h[128,256]", 1, 1, 256, 256, 128, 2)]
for label, a_mn, b_mn, m, n, k, kind in shapes:
    for split in (0, 1, 2, 8):
        ms = C.c_float()
        st = (C.c_int64 * 8)()
        rc = lib.ocf_gemm_tc_profile(a_mn, b_mn, m, n, k, split, kind, 50, C.byref(ms), st)
        if rc != 0:
            print(label, "split", split, "->", _lib.last_error() if hasattr(_lib, "last_error") else rc)
            continue
        print("%-44s split %d: %6.2f us/launch | %s" % (label, split, 1e3 * ms.value,
              "  ".join("%s %.1f" % (nm, v / 1e3) for nm, v in zip(names[1:], list(st)[1:]))))
