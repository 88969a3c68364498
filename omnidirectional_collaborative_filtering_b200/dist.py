"""Multi-GPU, one process per GPU: column (catalogue-dimension) sharding, and row (data)
parallelism with a gradient all-reduce.

Why columns. At the reference's batch size the step is bound by the weight + optimizer-state
stream, not by the batch, so replicating the weights and all-reducing
their gradients (66-440 MB per step on ML-10M shapes) cannot scale. The model is separable over
catalogue columns instead: the encoder is a sum over input columns and the decoder, loss and both
big weight updates are per output column. Rank g of G therefore owns columns [gN/G, (g+1)N/G):
those rows of W_enc (every input block), those rows of W_dec^T / entries of b_dec, and those
ratings of every store. All ranks walk the SAME global batch of rows (same NumPy stream, same row
order, same keep flags), and one step exchanges only activations:

    phase 1  gather + encoder partial sums          -> all-reduce z        [rows, H]
    phase 2  activations, decoder, loss partials    -> all-reduce (row stats | dL/dh)  [rows, 4 + H]
    phase 3  backward + fused optimizer update of the rank's own columns, metrics

Hidden layers and biases are replicated and receive identical updates on every rank. The result
is the single-GPU model at global batch `rows` (up to fp32 summation order), which is how a
data-parallel run of G x per-GPU-batch is obtained here ("weak" scaling in bench.py).

Row parallelism (`row_parallel_model`, SURVEY.md section 8e first row) is the textbook scheme the
north star names: replicated weights, rank g steps on rows [gB, (g+1)B) of every global batch,
the gradients of all trainable parameters are summed with ONE NCCL all-reduce and applied by a
streaming optimizer pass. It moves the whole parameter set over NVLink every step (66-440 MB on
ML-10M shapes) and cannot fuse the update into the weight-gradient kernel, so it is provided for
completeness and for small models; column sharding is the path that scales.

The collectives of both schemes run INSIDE the C library (`ocf_comm_*`, `ocf_model_set_comm`:
ncclAllReduce / ncclAllGather on the step's stream, one C call per step). `torch.distributed`
only boots the ranks and ships the NCCL id; `ShardComm` (collectives on torch tensors aliasing the
library's buffers, gloo-capable) remains for the CPU tests of the host logic and for runs without
libnccl.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from . import _lib


def col_range(n_cols: int, rank: int, world: int):
    return rank * n_cols // world, (rank + 1) * n_cols // world


# ---- weights: full Keras lists <-> per-rank slices ------------------------------------------------
def slice_weights(full: Sequence[np.ndarray], k_blocks: int, n_cols: int, lo: int, hi: int) -> List[np.ndarray]:
    """The rank's part of a full `get_weights()` list: rows [blk*N+lo, blk*N+hi) of the first
    kernel for every input block, columns [lo, hi) of the last kernel and bias; the rest whole."""
    out = [np.array(w, dtype=np.float32) for w in full]
    out[0] = np.concatenate([full[0][b * n_cols + lo:b * n_cols + hi] for b in range(k_blocks)], axis=0)
    out[-2] = np.ascontiguousarray(full[-2][:, lo:hi])
    out[-1] = np.ascontiguousarray(full[-1][lo:hi])
    return out


def merge_weights(parts: Sequence[Sequence[np.ndarray]], k_blocks: int, n_cols: int) -> List[np.ndarray]:
    """Inverse of `slice_weights` over all ranks' lists (rank order)."""
    world = len(parts)
    out = [np.array(w) for w in parts[0]]
    blocks = []
    for b in range(k_blocks):
        for r in range(world):
            lo, hi = col_range(n_cols, r, world)
            w = parts[r][0]
            blocks.append(w[b * (hi - lo):(b + 1) * (hi - lo)])
    out[0] = np.concatenate(blocks, axis=0)
    out[-2] = np.concatenate([p[-2] for p in parts], axis=1)
    out[-1] = np.concatenate([p[-1] for p in parts], axis=0)
    return out


class _DeviceArray(object):
    """Minimal `__cuda_array_interface__` carrier so torch can alias a library buffer."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class ShardComm(object):
    """The two collectives of a sharded step, on torch tensors aliasing the model's buffers."""

    def __init__(self, net, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.net = net
        self._capacity = None
        self.z = self.stats_dh = None

    def _alias(self):
        net = self.net
        if self._capacity == net._capacity and self.z is not None:
            return
        lib = _lib.lib()
        ptr, cnt = C.c_void_p(), C.c_int64()
        _lib.check(lib.ocf_model_buffer(net._handle, _lib.BUF_Z, C.byref(ptr), C.byref(cnt)))
        self.z = self.torch.as_tensor(_DeviceArray(ptr.value, cnt.value), device="cuda")
        _lib.check(lib.ocf_model_buffer(net._handle, _lib.BUF_STATS_DH, C.byref(ptr), C.byref(cnt)))
        self.stats_dh = self.torch.as_tensor(_DeviceArray(ptr.value, cnt.value), device="cuda")
        self._capacity = net._capacity
        self.hp0 = self.z.numel() // net._capacity[0]
        self.hpt = self.stats_dh.numel() // net._capacity[0] - 4

    def reduce_z(self, rows: int):
        self._alias()
        self.dist.all_reduce(self.z[:rows * self.hp0], group=self.group)

    def reduce_stats_dh(self, rows: int, with_dh: bool):
        self._alias()
        n = 4 * self._capacity[0] + (rows * self.hpt if with_dh else 0)
        self.dist.all_reduce(self.stats_dh[:n], group=self.group)


class NativeComm(object):
    """This rank's NCCL communicator inside the C library (`ocf_comm_create`). The 128-byte NCCL
    id is drawn on rank 0 and shipped over the already initialised torch.distributed group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        lib = _lib.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ident = (C.c_uint8 * 128)()
        if self.rank == 0:
            _lib.check(lib.ocf_comm_unique_id(ident))
        box = [bytes(ident) if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
        out = C.c_void_p()
        _lib.check(lib.ocf_comm_create(ident, self.rank, self.world, C.byref(out)))
        self.handle = out
        info = (C.c_int32 * 3)()
        _lib.check(lib.ocf_comm_info(out, info))

    def close(self):
        if self.handle is not None:
            _lib.lib().ocf_comm_destroy(self.handle)
            self.handle = None


def sharded_model(rank: int, world: int, numlayers, num_hidden_units, input_shape, batch_size, native=None, **kw):
    """`omni_model` for this rank's column slice. Every rank must call it with the same NumPy
    global RNG state: the full glorot initialisation is drawn identically everywhere and sliced.
    `native`: a NativeComm -> the step's two all-reduces run inside the C library."""
    from .model import omni_model
    lo, hi = col_range(int(input_shape), rank, world)
    om = omni_model(numlayers, num_hidden_units, input_shape, batch_size, local_cols=hi - lo, col_lo=lo, **kw)
    om.model.comm = ShardComm(om.model)
    if native is not None:
        om.model.native = (native, _lib.PAR_COLUMNS)
    return om


def row_parallel_model(native, numlayers, num_hidden_units, input_shape, batch_size, **kw):
    """Replicated `omni_model` for data parallelism over rows (same NumPy stream on every rank ->
    same initialisation). Feed it `batch.row_slice(rank, world)` of every global batch. `native`
    None = a single rank (the gradient path without the all-reduce)."""
    from .model import omni_model
    om = omni_model(numlayers, num_hidden_units, input_shape, batch_size, **kw)
    om.model.native = (native, _lib.PAR_ROWS)
    return om


def merge_topk(parts: Sequence, k: int):
    """The k best columns per row over the whole catalogue from the shards' own top-k lists
    (`model.recommend` of a column shard returns global column ids): (columns int32 [B, k], scores float32 [B, k]),
    best first, ties broken by the lower column - the order `recommend` itself uses. Empty slots (column -1,
    score -inf) sort last."""
    cols = np.concatenate([np.asarray(c, dtype=np.int32) for c, _ in parts], axis=1)
    scores = np.concatenate([np.asarray(v, dtype=np.float32) for _, v in parts], axis=1)
    empty = cols < 0
    # lexsort: last key is primary. Primary: empty slots last; then score descending; then column ascending
    order = np.lexsort((cols, -scores.astype(np.float64), empty), axis=1)[:, :int(k)]
    rows = np.arange(cols.shape[0])[:, None]
    out_c, out_s = cols[rows, order], scores[rows, order]
    if out_c.shape[1] < k:                           # fewer candidates than k: pad like recommend does
        pad = int(k) - out_c.shape[1]
        out_c = np.concatenate([out_c, np.full((out_c.shape[0], pad), -1, dtype=np.int32)], axis=1)
        out_s = np.concatenate([out_s, np.full((out_s.shape[0], pad), -np.inf, dtype=np.float32)], axis=1)
    return out_c, out_s


def all_gather_topk(cols, scores, k: int, group=None):
    """`merge_topk` over the ranks of a column-sharded model: every rank gets the catalogue-wide top k."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    packed = torch.from_numpy(np.concatenate([np.asarray(cols, dtype=np.int32).view(np.float32),
                                              np.asarray(scores, dtype=np.float32)], axis=1).copy())
    if dist.get_backend(group) == "nccl":
        packed = packed.cuda()
    bucket = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(bucket, packed, group=group)
    kk = np.asarray(cols).shape[1]
    parts = []
    for t in bucket:
        a = t.cpu().numpy()
        parts.append((a[:, :kk].copy().view(np.int32), a[:, kk:]))
    return merge_topk(parts, k)


def gather_full_weights(om, group=None) -> List[np.ndarray]:
    """Full Keras-layout weight list assembled from all ranks (every rank gets it)."""
    import torch.distributed as dist
    local = om.model.get_weights()
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    return merge_weights(parts, om.k_blocks, om.input_shape)


# ---- bench.py, N > 1 ----------------------------------------------------------------------------------
_KERNEL_TAGS = {0: "k_gather_split (K1)", 1: "k_enc_fwd (K2)", 2: "k_dec_fwd (K3)", 3: "k_sort_count+alloc+place (K4a)", 5: "k_row_update (K4b)"}


def _model_kwargs(w):
    aux = w["aux"]
    return dict(dense_activation=w["act"], use_causal_info=aux is not None, use_both_masks=aux == "both",
                dropout_probability=w["dropout"], auxilliary_mask_type=aux)


def _parity_vs_one_rank(args, w, fs, rd, om, B, rank, world):
    """The first two steps of the sharded model (fresh weights) next to the same two steps of an unsharded model on
    rank 0 (same initialisation, same NumPy stream, same global batches): max relative difference of the six step
    metrics. Collective: every rank steps its shard; only rank 0 also builds and steps the 1-rank model."""
    import contextlib
    import sys
    import torch.distributed as dist
    from . import optimizers
    from .data_reader import data_reader, sync_host_rng
    from .model import omni_model
    aux = w["aux"]

    def two_steps(reader, model):
        sync_host_rng()
        np.random.seed(11)
        g = reader.data_gen(B, w["sparsity"], "train", True, aux, w["aux_value"], pass_through_input_training=w["pass_through"])
        out = [model.train_on_batch(next(g), sync=True) for _ in range(2)]
        sync_host_rng()
        return np.array(out, dtype=np.float64)

    got = two_steps(rd, om.model)
    res = None
    if rank == 0:
        with contextlib.redirect_stdout(sys.stderr):
            rd1 = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
        np.random.seed(0)
        om1 = omni_model(w["layers"], w["hidden"], fs.n_cols, B, **_model_kwargs(w))
        opt = {"adagrad": optimizers.Adagrad, "rmsprop": optimizers.RMSprop, "adam": optimizers.Adam}[w["opt"][0]](lr=w["opt"][1])
        om1.model.compile(opt, "mean_squared_error", rating_range=fs.rating_range)
        want = two_steps(rd1, om1.model)
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-12)
        res = {"steps": 2, "max_rel_diff_of_step_metrics": float(rel.max()), "bar": 1e-3, "ok": bool(rel.max() < 1e-3),
               "sharded_accurate_RMSE": [float(v) for v in got[:, 3]], "one_rank_accurate_RMSE": [float(v) for v in want[:, 3]],
               "what": "first two train steps of the %d-rank column-sharded model vs the unsharded model on rank 0 "
                       "(same init, same NumPy stream, global batch %d rows)" % (world, B)}
        om1.model.close(); rd1.close()
    dist.barrier()
    return res


def _bench_config(args, w, rank, world, native, B, scaling, parity=False, steps=None):
    """One (workload, global batch) configuration on all ranks: device-timed value (CUDA events, barrier + synchronize
    on both sides, max over ranks), per-rank kernel / exchange times, e2e through the public API. Returns the record on
    rank 0 (None elsewhere); every rank makes the same collective calls."""
    import contextlib
    import sys
    import time
    import torch
    import torch.distributed as dist
    import bench
    from . import optimizers
    from .data_reader import data_reader
    from .store import DeviceBatch

    lib = _lib.lib()
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    rows_mode = getattr(args, "parallel", "columns") == "rows"
    fs = None
    if rank == 0:
        fs = bench.make_dataset(w)             # generate (or load the /tmp cache) once
    dist.barrier()
    if fs is None:
        fs = bench.make_dataset(w)
    with contextlib.redirect_stdout(sys.stderr):     # the reader prints the reference's "Finished loading data"
        rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs,
                         shard=None if rows_mode else (rank, world))
    aux = w["aux"]
    np.random.seed(0)
    if rows_mode:
        om = row_parallel_model(native, w["layers"], w["hidden"], fs.n_cols, B // world, **_model_kwargs(w))
    else:
        om = sharded_model(rank, world, w["layers"], w["hidden"], fs.n_cols, B, native=native, **_model_kwargs(w))
    m = om.model
    opt = {"adagrad": optimizers.Adagrad, "rmsprop": optimizers.RMSprop, "adam": optimizers.Adam}[w["opt"][0]](lr=w["opt"][1])
    m.compile(opt, "mean_squared_error", rating_range=fs.rating_range)
    K, W = int(steps or args.steps), max(args.warmup, 3)
    K = max(2, min(K, rd.train_set_size // B * 4))
    parity_rec = None
    if parity and not rows_mode:
        try:
            parity_rec = _parity_vs_one_rank(args, w, fs, rd, om, B, rank, world)
        except Exception as exc:                         # pragma: no cover - must not take the line down
            parity_rec = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
            dist.barrier()

    def gen():
        return rd.data_gen(B, w["sparsity"], "train", True, aux, w["aux_value"], pass_through_input_training=w["pass_through"])

    def endless():
        g = gen()
        while True:
            b = next(g)
            if b is None:
                g = gen()
                continue
            yield b.row_slice(rank, world) if rows_mode else b

    batches = endless()
    plans = []
    for _ in range(K + W):
        p = next(batches)
        p.flags                             # spend the batch's draws before the next one is drawn
        plans.append(p)
    m._ensure(plans[0].n_rows, max(p.n_entries for p in plans), aux, rd)
    resident = []
    for p in plans:
        dev = DeviceBatch(p.n_rows, p.n_entries)
        dev.fill_split(p.source, p.rows, p.flags, p.pass_through, p.aux_value, None)
        resident.append(dev)

    def device_steps(devs, first):
        for k, dev in enumerate(devs):
            _lib.check(lib.ocf_batch_regather(dev.handle, None))
            m.step_on_device_batch(dev, B // world if rows_mode else B, first + k, train=True,
                                   row0=rank * (B // world) if rows_mode else 0, rows_total=B if rows_mode else 0)

    # two passes over every resident batch object before the clock starts (plain launches, then graph capture)
    device_steps(resident, 0)
    device_steps(resident, K + W)
    device_steps(resident[:W], 2 * (K + W))
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.ocf_kernel_launches()
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    device_steps(resident[W:], 2 * (K + W) + W)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = lib.ocf_kernel_launches() - launches0
    own_ms = e0.elapsed_time(e1)
    t = torch.tensor([own_ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # keep the same steps running ~0.6 s so nvidia-smi (100 ms period) samples clocks under this
    # load; the repeat count comes from the all-reduced time, so every rank runs the same number
    # of collectives
    for _ in range(max(1, min(200, int(600.0 / max(ms, 1.0))))):
        device_steps(resident[W:], W)
    torch.cuda.synchronize()
    dist.barrier()
    # per-kernel / per-collective CUDA-event times of one more pass, on EVERY rank (plain launches: the
    # instrumented pass does not replay the graphs)
    lib.ocf_profile_reset()
    lib.ocf_profile_enable(1)
    device_steps(resident[W:], W)
    torch.cuda.synchronize()
    lib.ocf_profile_enable(0)
    dist.barrier()
    names = dict(_KERNEL_TAGS)
    names[6] = "streaming optimizer pass" if rows_mode else "exchange z [rows, H] (ncclAllReduce)"
    names[7] = "ncclAllReduce gradients" if rows_mode else "exchange row stats + dL/dh [rows, 4 + H]"
    kernels = {}
    for tag, name in names.items():
        tot, cnt = C.c_double(), C.c_int64()
        _lib.check(lib.ocf_profile_read(tag, C.byref(tot), C.byref(cnt)))
        kernels[name] = {"ms": tot.value / max(cnt.value, 1)}
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {"rank": rank, "own_timed_ms_per_step": own_ms / K,
                                      "ms": {k: round(v["ms"], 5) for k, v in kernels.items()}})
    # roofline of rank 0's dominant kernel (the fused row update): algorithmic bytes of the rank's own
    # shard / CUDA-event time; never allowed to break the line
    roofline = None
    state_mb = sum(float(np.prod(sh)) for sh in om.weight_shapes()) * 4 * bench.state_words_of(w) / 2 / 1e6
    if not rows_mode:
        try:
            alg = bench.step_bytes(plans[W:], w, 0)
            k4b_ms = kernels["k_row_update (K4b)"]["ms"]
            peak, peak_src = bench.peaks()
            if k4b_ms > 0:
                ach = alg[4] / (k4b_ms * 1e-3) / 1e9
                roofline = {"kernel": "k_row_update (K4b), rank 0's shard", "bound": "hbm" if state_mb > 2 * 126 else "l2",
                            "achieved": ach, "peak": peak,
                            "unit": "GB/s", "frac": ach / peak, "traffic": None, "algorithmic_bytes": alg[4],
                            "peak_source": peak_src, "share_of_step": k4b_ms / (ms / K)}
        except Exception as exc:                      # pragma: no cover
            roofline = {"error": str(exc)}
    rt = torch.tensor([float(sum(p.n_ratings for p in plans[W:])),
                       float(sum(p.n_ratings if p.pass_through else 0 for p in plans[W:]))], device="cuda")
    if rows_mode:
        dist.all_reduce(rt)                              # every rank holds its own rows of the global batches
    ratings = float(rt[0].item())                        # ratings of the global batches (all ranks together)
    # SURVEY 8(d)'s rating = one observed TARGET entry: each rank counts the targets among its own ratings
    tg = torch.tensor([float(sum(p.n_entries if p.pass_through else int((np.asarray(p.flags) == 0).sum()) for p in plans[W:]))],
                      device="cuda")
    dist.all_reduce(tg)
    targets = float(tg.item())

    # e2e through the public API on every rank: generator thread (row ids, the batch's place in the NumPy stream),
    # H2D of the row ids, device-side draw of the random split, phases / collectives, D2H of the step's metrics
    from .data_reader import Prefetcher
    per_epoch = max(rd.train_set_size // B, 1)

    def epochs(total):
        """`total` batches epoch by epoch like train.py:150-158 (a fresh generator per epoch), drawn on a generator
        thread like Keras' GeneratorEnqueuer; every rank draws the same batches in the same order."""
        done = 0
        while done < total:
            n = min(total - done, per_epoch)
            for bt in Prefetcher(gen(), n):
                yield bt.row_slice(rank, world) if rows_mode else bt
            done += n

    for b in epochs(max(W, 9)):              # three rounds of the ring of 3 batch buffers: plain, capture, replay
        m.train_on_batch(b, sync=True)
    dist.barrier()
    torch.cuda.synchronize()
    h2d = e_ratings = 0
    from collections import deque
    in_flight = deque()
    lag = max(1, int(os.environ.get("OCF_BENCH_LAG", "2")))      # like bench.py's 1-GPU loop: every step's metrics are read, `lag` steps behind
    t0 = time.perf_counter()
    for b in epochs(K):
        m.train_on_batch(b, sync=False)
        in_flight.append(m.steps_logged() - 1)
        if len(in_flight) > lag:
            m.wait_metrics(in_flight.popleft())
        e_ratings += b.n_ratings
        h2d += b._device.info()["h2d_bytes"]
    while in_flight:
        m.wait_metrics(in_flight.popleft())
    torch.cuda.synchronize()
    dist.barrier()
    e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    if rows_mode:
        er = torch.tensor([float(e_ratings)], device="cuda")
        dist.all_reduce(er)
        e_ratings = float(er.item())
    h2d_all = torch.tensor([float(h2d)], device="cuda")
    dist.all_reduce(h2d_all)
    clocks = sampler.stop() if sampler else None
    rec = None
    if rank == 0:
        if rows_mode:
            par = ("row-parallel x%d, global batch %d rows (%d per GPU), replicated weights, one NCCL all-reduce of all gradients per step"
                   % (world, B, B // world))
        else:
            par = ("column-sharded x%d, global batch %d rows (every rank walks all of them on its own columns), 2 exchanges of [rows, H] per step as ncclAllReduce (captured in the step's CUDA graph)"
                   % (world, B))
        l2 = ("no flush: per-rank weights + optimizer state (%.0f MB) exceed the 126 MB L2 several times over" % state_mb
              if state_mb > 2 * 126 else
              "no flush, and per-rank weights + optimizer state (%.0f MB) fit or nearly fit the 126 MB L2: the rank's kernels "
              "run from L2, an HBM fraction is not meaningful here (roofline.bound says 'l2')" % state_mb)
        rec = {"value": ratings / (ms * 1e-3), "unit": "ratings/s", "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms / K, "scaling": scaling, "global_batch_rows": B, "parallelism": par, "l2": l2,
               "ratings_per_step": ratings / K, "target_ratings_per_step": targets / K,
               "target_ratings_per_s": targets / (ms * 1e-3), "clocks": clocks,
               "e2e": {"value": e_ratings / float(e2e.item()), "unit": "ratings/s",
                       "h2d_bytes_per_step": float(h2d_all.item()) / K, "d2h_bytes_per_step": 4 * _lib.N_METRICS * world,
                       "ms_per_step": 1e3 * float(e2e.item()) / K},
               "gpu_launches": int(launches), "roofline": roofline, "kernels_rank0": kernels, "kernels_per_rank": per_rank}
        if parity_rec is not None:
            rec["parity"] = parity_rec
    del resident
    m.close()
    rd.close()
    dist.barrier()
    return rec


def bench_main(args, w, cfg, rank, world):
    """`bench.py --gpus N` under torchrun, columns sharded over the ranks. The line's `value` is weak scaling (global
    batch = N x batch_size rows: the data-parallel run of N per-GPU batches); `strong_scaling` carries the same model at
    the reference's own global batch (batch_size rows in all, train.py:30); at N = 8 the Netflix-shaped config rides
    along under `other_workloads` (it is the config the 8-way column split exists for). Rank 0 prints the JSON line."""
    import json
    import torch
    import torch.distributed as dist
    import bench

    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    native = NativeComm()
    rows_mode = getattr(args, "parallel", "columns") == "rows"
    main = _bench_config(args, w, rank, world, native, args.batch_size * world, "weak", parity=True)
    strong = None
    if not rows_mode and os.environ.get("OCF_BENCH_STRONG", "1") != "0":
        strong = _bench_config(args, w, rank, world, native, args.batch_size, "strong", steps=min(args.steps, 30))
    others = getattr(args, "others", None)
    if others is None:
        others = "netflix" if (world == 8 and args.workload == "ml10m" and not rows_mode) else "none"
    other_recs = {}
    for name in [x for x in others.split(",") if x and x != "none"]:
        try:
            other_recs[name] = _bench_config(args, bench.WORKLOADS[name], rank, world, native, args.batch_size * world, "weak",
                                             parity=True, steps=min(args.steps, 12))
        except Exception as exc:                          # pragma: no cover
            other_recs[name] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    if rank == 0:
        line = {"metric": "train ratings/sec", "value": main["value"], "unit": "ratings/s", "n_gpus": world,
                "steps": main["steps"], "warmup": main["warmup"], "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg}
        for k, v in main.items():
            if k not in line:
                line[k] = v
        if strong is not None:
            line["strong_scaling"] = {k: strong[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "global_batch_rows",
                                                             "ratings_per_step", "e2e", "gpu_launches", "kernels_rank0", "l2")}
        if other_recs:
            line["other_workloads"] = {k: ({kk: vv for kk, vv in v.items() if kk != "kernels_per_rank"} if "error" not in v else v)
                                       for k, v in other_recs.items()}
            for k, v in other_recs.items():
                if "error" not in v:
                    line["other_workloads"][k]["kernels_per_rank"] = v["kernels_per_rank"]
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0
