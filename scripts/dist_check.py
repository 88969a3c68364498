"""Real multi-GPU check (run under torchrun, one rank per GPU): the column-sharded step and the
row-parallel (gradient all-reduce) step, both with their collectives inside the C library, must
reproduce the unsharded oracle at the global batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_check.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from oracle import ref_batches, ref_model
from omnidirectional_collaborative_filtering_b200 import dist as ocf_dist, optimizers, synthetic
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader


def check(mode, native, rank, world):
    fs = synthetic.make_fixed_split("small", reverse_user_item_data=True, seed=4)
    B = 32 * world
    rows_mode = mode == "rows"
    rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs,
                     shard=None if rows_mode else (rank, world))
    np.random.seed(0)
    kw = dict(dense_activation="tanh", use_causal_info=True, dropout_probability=0.2, auxilliary_mask_type="dropout")
    if rows_mode:
        om = ocf_dist.row_parallel_model(native, 2, 96, fs.n_cols, B // world, **kw)
    else:
        om = ocf_dist.sharded_model(rank, world, 2, 96, fs.n_cols, B, native=native, **kw)
    om.model.compile(optimizers.Adagrad(lr=0.01), "mean_squared_error", rating_range=4.0)
    full0 = om.model.get_weights() if rows_mode else ocf_dist.gather_full_weights(om)
    ref = ref_model.RefModel(2, 96, fs.n_cols, B, dense_activation="tanh", use_causal_info=True,
                             dropout_probability=0.2, dtype=np.float32)
    ref.set_weights(full0)
    ref.dropout_seed = om.dropout_seed
    ref.compile(ref_model.RefOptimizer("adagrad", lr=0.01), "mean_squared_error", rating_range=4.0)
    dicts = synthetic.to_reference_dicts(fs, raw_col_id=lambda c: c)
    data = ref_batches.RefData(fs.n_cols, fs.train.n_rows, dicts["unique_cols"], eval_mode="fixed_split",
                               train=dicts["train"], valid=tuple(dicts["valid"]), test=tuple(dicts["test"]))
    np.random.seed(7)
    gen = rd.data_gen(B, [0.4, 0.9], "train", True, "dropout", -1)
    rng7 = np.random.RandomState(7)
    rgen = ref_batches.batch_stream(data, B, [0.4, 0.9], "train", True, "dropout", -1, rng=rng7, vectorised=True)
    worst = 0.0
    for step in range(9):          # ring of 3 batch buffers: plain launches, captured steps (NCCL inside the graph), replays
        b, rb = next(gen), next(rgen)
        if b is None:                  # the epoch is over (on both sides): a new generator each, like train.py:153
            assert rb is None
            gen = rd.data_gen(B, [0.4, 0.9], "train", True, "dropout", -1)
            rgen = ref_batches.batch_stream(data, B, [0.4, 0.9], "train", True, "dropout", -1, rng=rng7, vectorised=True)
            b, rb = next(gen), next(rgen)
        got = om.model.train_on_batch(b.row_slice(rank, world) if rows_mode else b)
        feed, targets = rb
        want = ref.train_on_batch(feed, targets)
        worst = max(worst, float(np.max(np.abs(np.array(got) - np.array(want)) / np.maximum(np.abs(want), 1e-6))))
    full = om.model.get_weights() if rows_mode else ocf_dist.gather_full_weights(om)
    wdiff = max(float(np.max(np.abs(a - b))) for a, b in zip(full, ref.get_weights()))
    ok = worst < 1e-3 and wdiff < 0.03
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist_check %s world=%d: max rel metric diff %.2e, max weight diff %.2e -> %s"
              % (mode, world, worst, wdiff, "OK" if flag.item() == 1.0 else "FAIL"))
    rd.close()
    return flag.item() == 1.0


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
    native = ocf_dist.NativeComm()
    ok = all([check(mode, native, rank, world) for mode in ("columns", "rows")])
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
