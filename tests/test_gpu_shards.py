"""GPU parity of the column-sharded step. Two shards run as two models on ONE GPU; the test plays
the role of the collectives by summing their exchange buffers between phases (the same three
phases dist.py drives with NCCL), and the result must match the unsharded oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ref_batches, ref_model
from omnidirectional_collaborative_filtering_b200 import _lib, dist as ocf_dist, optimizers
from tests.helpers import oracle_data
from tests.test_shard_host import _shard_reader

pytestmark = pytest.mark.gpu


class _LocalSum(object):
    """Stands in for NCCL: all shards live in this process, reductions are plain tensor sums."""

    def __init__(self, comms):
        self.comms = comms

    def reduce(self, name, n):
        for c in self.comms:
            c._alias()
        bufs = [getattr(c, name)[:n] for c in self.comms]
        total = torch.stack(bufs).sum(dim=0)
        for b in bufs:
            b.copy_(total)


@pytest.mark.parametrize("aux,layers,opt,pdrop", [(None, 1, "adagrad", 0.2), ("both", 2, "adam", None),
                                                 ("causal", 1, "rmsprop", 0.3)])
def test_two_shards_match_unsharded_oracle(golden_datasets, aux, layers, opt, pdrop):
    ds = golden_datasets["rev"]
    N, B, world = ds["n_cols"], 8, 2
    kw = dict(dense_activation="sigmoid", use_causal_info=aux is not None, use_both_masks=aux == "both",
              dropout_probability=pdrop)
    readers = [_shard_reader(ds, "fixed_split", r, world) for r in range(world)]
    models = []
    for r in range(world):
        np.random.seed(11)                                 # identical init stream on every rank
        om = ocf_dist.sharded_model(r, world, layers, 40, N, B, auxilliary_mask_type=aux, **kw)
        models.append(om)
    o = {"adagrad": optimizers.Adagrad(lr=0.05), "adam": optimizers.Adam(lr=0.01), "rmsprop": optimizers.RMSprop(lr=0.01)}[opt]
    for om in models:
        om.model.compile(o, "mean_squared_error", rating_range=4.0)
    np.random.seed(11)
    from omnidirectional_collaborative_filtering_b200.model import omni_model
    full = omni_model(layers, 40, N, B, auxilliary_mask_type=aux, **kw)       # same stream -> same full init
    ref = ref_model.RefModel(layers, 40, N, B, dtype=np.float32, rng=np.random.RandomState(0), **kw)
    ref.set_weights(full.model.get_weights())
    ref.dropout_seed = full.dropout_seed
    ref.compile(ref_model.RefOptimizer(opt, lr=o.lr), "mean_squared_error", rating_range=4.0)
    merged0 = ocf_dist.merge_weights([om.model.get_weights() for om in models], full.k_blocks, N)
    for a, b in zip(merged0, ref.get_weights()):
        assert np.array_equal(a, b)

    data = oracle_data(ds, "fixed_split")
    rgen = ref_batches.batch_stream(data, B, [0.3, 0.8], "train", True, aux, -1, rng=np.random.RandomState(3))
    gens = []
    for rd in readers:
        np.random.seed(3)
        g = rd.data_gen(B, [0.3, 0.8], "train", True, aux, -1)
        gens.append([next(g) for _ in range(4)])           # drain each rank's stream separately
        for b in gens[-1]:
            b.flags        # one process plays both ranks here: spend rank r's draws before the stream is reseeded
    lib = _lib.lib()
    summer = _LocalSum([om.model.comm for om in models])
    for step in range(4):
        handles, devs = [], []
        for om, batches in zip(models, gens):
            b = batches[step]
            h = om.model._ensure(b.n_rows, b.n_entries, b.aux_type, b.reader)
            handles.append(h); devs.append(b.upload(None))
        recs = [np.empty(_lib.N_METRICS, dtype=np.float32) for _ in models]
        for phase in (1, 2, 3):
            for om, h, dev, rec in zip(models, handles, devs, recs):
                args = om.model._args(None, phase=phase)
                args.step = step
                _lib.check(lib.ocf_train_step(h, dev.handle, C.byref(args), _lib.ptr(rec) if phase == 3 else None, None))
            c = models[0].model.comm
            c._alias()
            if phase == 1:
                summer.reduce("z", B * c.hp0)
            elif phase == 2:
                summer.reduce("stats_dh", 4 * c._capacity[0] + B * c.hpt)
        feed, targets = next(rgen)
        want = ref.train_on_batch(feed, targets)
        for rec in recs:                                   # every rank logs the same global metrics
            np.testing.assert_allclose(rec[:6], want, rtol=3e-4, atol=1e-6)
    merged = ocf_dist.merge_weights([om.model.get_weights() for om in models], full.k_blocks, N)
    for g, w in zip(merged, ref.get_weights()):
        bad = np.abs(g - w) > (2e-6 + 1e-3 * np.abs(w))
        assert bad.mean() < 1e-3 and np.max(np.abs(g - w)) <= 3 * o.lr
    # replicated parameters stay identical across shards
    w0, w1 = models[0].model.get_weights(), models[1].model.get_weights()
    for i in range(1, len(w0) - 2):
        assert np.array_equal(w0[i], w1[i])
