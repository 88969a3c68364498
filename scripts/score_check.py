#!/usr/bin/env python
"""Full-catalogue scoring on one B200: the tcgen05 kernel's time (CUDA events around it on the
launching stream), TFLOP/s and output GB/s per batch size, plus a large-shape consistency check
against the fp32 SDDMM path (`predict` = mask * full at the target entries).

    python scripts/score_check.py [--workload ml10m] [--rows 128,1024,4096]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ml10m")
    ap.add_argument("--rows", default="128,1024,4096")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    import bench
    from omnidirectional_collaborative_filtering_b200 import _lib
    from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
    from omnidirectional_collaborative_filtering_b200.model import omni_model

    w = bench.WORKLOADS[args.workload]
    fs = bench.make_dataset(w)
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
    aux = w["aux"]
    lib = _lib.lib()
    N = fs.n_cols
    H = w["hidden"] if isinstance(w["hidden"], int) else w["hidden"][-1]
    hp = (H + 127) // 128 * 128
    out_lines = []
    for B in [int(x) for x in args.rows.split(",")]:
        B = min(B, rd.val_set_size)
        np.random.seed(0)
        om = omni_model(w["layers"], w["hidden"], N, B, dense_activation=w["act"], use_causal_info=aux is not None,
                        use_both_masks=aux == "both", auxilliary_mask_type=aux)
        m = om.model
        wts = m.get_weights()
        rs = np.random.RandomState(1)
        wts[-1] = rs.normal(size=wts[-1].shape).astype(np.float32)
        wts[-2] = (rs.normal(size=wts[-2].shape) * 0.2).astype(np.float32)
        m.set_weights(wts)
        np.random.seed(2)
        batch = next(rd.data_gen(B, None, "valid", True, aux, w["aux_value"]))
        h = m._ensure(batch.n_rows, batch.n_entries, batch.aux_type, batch.reader)
        dev = batch.upload(m.stream)
        out = torch.empty((B, N), dtype=torch.float32, device="cuda")
        _lib.check(lib.ocf_score(h, dev.handle, C.c_void_p(out.data_ptr()), 1, m.stream))
        torch.cuda.synchronize()
        # consistency with the fp32 SDDMM path at the target entries
        pred = m.predict(batch)                       # mask * full, zeros elsewhere
        got = out.cpu().numpy()
        sel = pred != 0
        full_ref = pred[sel] / w["aux_value"]
        err = got[sel] - full_ref
        lib.ocf_profile_reset(); lib.ocf_profile_enable(1)
        for _ in range(args.reps):
            _lib.check(lib.ocf_score(h, dev.handle, C.c_void_p(out.data_ptr()), 1, m.stream))
        torch.cuda.synchronize()
        lib.ocf_profile_enable(0)
        tot, cnt = C.c_double(), C.c_int64()
        _lib.check(lib.ocf_profile_read(4, C.byref(tot), C.byref(cnt)))
        ms = tot.value / max(cnt.value, 1)
        flops = 2.0 * B * hp * N
        line = {"workload": args.workload, "rows": B, "n_cols": N, "hp": hp, "kernel_ms": ms,
                "tflops": flops / (ms * 1e-3) / 1e12, "out_GBs": 4.0 * B * N / (ms * 1e-3) / 1e9,
                "rows_per_s": B / (ms * 1e-3), "checked_entries": int(sel.sum()),
                "max_abs_err_vs_fp32": float(np.abs(err).max()) if err.size else None,
                "mean_err": float(err.mean()) if err.size else None,
                "rms_ref": float(np.sqrt((full_ref ** 2).mean())) if err.size else None}
        print(json.dumps(line), flush=True)
        out_lines.append(line)
        m.close()
    rd.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
