"""Per-step host timeline of bench.py's e2e loop (debug helper): where do the microseconds between the device-timed
step and the end-to-end step go?  python scripts/e2e_trace.py [workload] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from collections import deque
from omnidirectional_collaborative_filtering_b200 import optimizers
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader, Prefetcher
from omnidirectional_collaborative_filtering_b200.model import omni_model

name = sys.argv[1] if len(sys.argv) > 1 else "ml10m"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
w = bench.WORKLOADS[name]
fs = bench.make_dataset(w)
rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
aux = w["aux"]
np.random.seed(0)
om = omni_model(w["layers"], w["hidden"], fs.n_cols, 128, dense_activation=w["act"], use_causal_info=aux is not None,
                use_both_masks=aux == "both", dropout_probability=w["dropout"], auxilliary_mask_type=aux)
m = om.model
m.compile({"adagrad": optimizers.Adagrad, "rmsprop": optimizers.RMSprop, "adam": optimizers.Adam}[w["opt"][0]](lr=w["opt"][1]),
          "mean_squared_error", rating_range=fs.rating_range)


def gen():
    return rd.data_gen(128, w["sparsity"], "train", True, aux, w["aux_value"], pass_through_input_training=w["pass_through"])


K = min(K, rd.train_set_size // 128)          # one epoch at most: a generator ends with its set
for _ in range(3):
    for b in Prefetcher(gen(), min(4, K)):
        m.train_on_batch(b, sync=True)
torch.cuda.synchronize()
for rep in range(2):
    stamps = []
    in_flight = deque()
    t0 = time.perf_counter()
    for b in Prefetcher(gen(), K):
        ta = time.perf_counter()
        m.train_on_batch(b, sync=False)
        tb = time.perf_counter()
        in_flight.append(m.steps_logged() - 1)
        if len(in_flight) > 2:
            m.wait_metrics(in_flight.popleft())
        tc = time.perf_counter()
        stamps.append((ta - t0, tb - ta, tc - tb))
        t_prev = tc
    while in_flight:
        m.wait_metrics(in_flight.popleft())
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print("rep %d: %d steps %.3f ms total = %.1f us/step" % (rep, K, 1e3 * total, 1e6 * total / K))
    arr = np.array(stamps)
    gaps = np.diff(arr[:, 0], prepend=0.0)
    for lo in range(0, K, 10):
        print("  steps %2d-%2d: start-to-start %s us | enqueue %s us | wait %s us" % (
            lo, lo + 9, np.round(1e6 * gaps[lo:lo + 10]).astype(int).tolist(), np.round(1e6 * arr[lo:lo + 10, 1]).astype(int).tolist(),
            np.round(1e6 * arr[lo:lo + 10, 2]).astype(int).tolist()))

# ---- anatomy of an epoch start: which piece of the ~1.5 ms is what ---------------------------------------------
import threading
from omnidirectional_collaborative_filtering_b200 import data_reader as dr_mod


def clock(label, fn):
    t = time.perf_counter()
    out = fn()
    print("  %-46s %7.0f us" % (label, 1e6 * (time.perf_counter() - t)))
    return out


for rep in range(2):
    print("epoch start, rep %d (device idle)" % rep)
    torch.cuda.synchronize()
    clock("sync_host_rng (stream back to np.random)", dr_mod.sync_host_rng)
    clock("np.random.permutation(%d)" % rd.train_set_size, lambda: np.random.permutation(rd.train_set_size))
    g = clock("data_gen(...) call", gen)
    b0 = clock("first next(generator)", lambda: next(g))
    b1 = clock("second next(generator)", lambda: next(g))
    clock("first train_on_batch (state to the device)", lambda: m.train_on_batch(b0, sync=False))
    clock("second train_on_batch", lambda: m.train_on_batch(b1, sync=False))
    clock("device drained", torch.cuda.synchronize)
    th = threading.Thread(target=lambda: None)
    clock("thread start + join", lambda: (th.start(), th.join()))
    it = iter(Prefetcher(gen(), 3))
    bb = clock("Prefetcher: first item", lambda: next(it))
    clock("Prefetcher: second item", lambda: next(it))
    for x in it:
        pass
