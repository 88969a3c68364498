"""Shared helpers for the test-suite (oracle-side construction of reference-shaped data)."""
import numpy as np

from oracle import ref_batches


def oracle_data(ds, eval_mode):
    """RefData over one of the golden datasets, wired like tests/golden/make_golden.py wires
    the reference reader."""
    if eval_mode == "ablation":
        return ref_batches.RefData(ds["n_cols"], len(ds["unique_rows"]), ds["unique_cols"],
                                   eval_mode="ablation", user_dict=ds["ablation"],
                                   nonsequentialusers=True, unique_rows=ds["unique_rows"])
    return ref_batches.RefData(ds["n_cols"], ds["n_rows"], ds["unique_cols"],
                               eval_mode="fixed_split", train=ds["train"],
                               valid=tuple(ds["valid"]), test=tuple(ds["test"]),
                               nonsequentialusers=True, unique_rows=ds["unique_rows"])


def golden_batch(npz, case, n):
    """(input list, targets, target_count or None) of batch n of a golden case."""
    cid = case["id"]
    feed = []
    k = 0
    while "%s/b%d/in%d" % (cid, n, k) in npz.files:
        feed.append(npz["%s/b%d/in%d" % (cid, n, k)])
        k += 1
    tc = npz["%s/b%d/target_count" % (cid, n)] if "%s/b%d/target_count" % (cid, n) in npz.files else None
    return feed, npz["%s/b%d/targets" % (cid, n)], tc


def product_reader(ds, eval_mode):
    """The product's data_reader over a golden dataset (in-memory 'files')."""
    from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
    files = {
        "unique_items_list": ds["unique_cols"], "unique_users_list": ds["unique_rows"],
        "ratingsByUser_dict": ds["ablation"],
        "ratingsByUser_dicts_train": ds["train"],
        "ratingsByUser_dicts_valid": ds["valid"],
        "ratingsByUser_dicts_test": ds["test"],
    }
    n_rows = len(ds["unique_rows"]) if eval_mode == "ablation" else ds["n_rows"]
    return data_reader(ds["n_cols"], n_rows, "", nonsequentialusers=True, use_json=True,
                       eval_mode=eval_mode, useTimestamps=False, reverse_user_item_data=False, data=files)


def host_densify(batch):
    """Dense arrays a product Batch stands for, computed on the host straight from its row ids and
    keep flags (test-side restatement of data_reader.py:158-169 / :234-268; no GPU involved)."""
    B, N, a = batch.n_rows, batch.n_cols, batch.aux_value
    mask_in, mask_out, x, t, observed = (np.zeros([B, N]) for _ in range(5))
    pos = 0
    for b, row in enumerate(batch.rows):
        if batch.kind == "split":
            cols, vals = batch.source.csr.row(int(row))
            f = batch.flags[pos:pos + len(cols)].astype(bool)
            pos += len(cols)
            for c, v, fi in zip(cols, vals, f):
                if fi:
                    mask_in[b, c] = a; x[b, c] = v
                    if batch.pass_through:
                        mask_out[b, c] = a; t[b, c] = v
                else:
                    mask_out[b, c] = a; t[b, c] = v
                observed[b, c] = a
        else:
            cols, vals = batch.source.in_store.csr.row(int(row))
            for c, v in zip(cols, vals):
                mask_in[b, c] = a; x[b, c] = v; observed[b, c] = a
            cols, vals = batch.source.tgt_store.csr.row(int(row))
            for c, v in zip(cols, vals):
                mask_out[b, c] = a; t[b, c] = v; observed[b, c] = a
    return ref_batches.feed_list((mask_in, mask_out, x, t, observed), batch.aux_type), t


class OracleNet(object):
    """The surface `train.run` uses of `omni_model.model`, computed by the NumPy oracle on host-densified
    product batches: lets the product's host logic (generators, RNG order, epoch loop, test procedures)
    run end to end without a GPU. Test infrastructure only."""

    _handle = None                       # what omni_model's helpers look at before talking to the library

    def __init__(self, ref, owner=None):
        from oracle import ref_model
        self.ref = ref
        self.owner = owner               # the product omni_model whose per-layer trainable flags apply
        self.metrics_names = list(ref_model.METRIC_NAMES)
        self.sse_log = []

    def compile(self, optimizer=None, loss="mean_squared_error", rating_range=1.0):
        from oracle import ref_model
        self.ref.compile(ref_model.RefOptimizer(optimizer.kind, lr=optimizer.lr, epsilon=optimizer.epsilon,
                                                decay=optimizer.decay), loss, rating_range=rating_range)

    def get_weights(self):
        return self.ref.get_weights()

    def set_weights(self, w):
        self.ref.set_weights(w)

    def fit_generator(self, gen, steps, validation_data=None, validation_steps=None, verbose=0):
        class _H(object):
            pass
        h = _H()
        if self.owner is not None:
            self.ref.trainable = list(self.owner.trainable)
        h.history = self.ref.fit_generator((host_densify(b) for b in gen), steps,
                                           validation_data=(host_densify(b) for b in validation_data),
                                           validation_steps=validation_steps)
        return h

    def evaluate_generator(self, gen, steps):
        return self.ref.evaluate_generator((host_densify(b) for b in gen), steps)

    def test_on_batch(self, batch, sync=True):
        feed, t = host_densify(batch)
        y = self.ref.predict(feed)
        self.sse_log.append(float(np.sum(np.square(np.subtract(y, t, dtype=np.float64)))))

    def steps_logged(self):
        return len(self.sse_log)

    def save(self, path):
        """The product's own `.npz` writer (it only needs `owner.config()` and `get_weights()`)."""
        from omnidirectional_collaborative_filtering_b200.model import OmniNet
        OmniNet.save(self, path)

    def read_metrics(self, first, count):
        rec = np.zeros((count, 8), dtype=np.float32)
        rec[:, 6] = self.sse_log[first:first + count]
        return rec


def oracle_train_run_ablation(user_dict, unique_cols, unique_rows, cfg, seed, init_model, rating_range):
    """The ablation branch of the script (train.py:85-86,147-177,202-213): rows split 80/10/10 by one permutation
    drawn BEFORE the model is built, every set goes through the random reciprocal split, the test set once per
    entry of `test_sparsities`."""
    from oracle import ref_model
    c = cfg
    data = ref_batches.RefData(len(unique_cols), len(unique_rows), unique_cols, eval_mode="ablation", user_dict=user_dict,
                               nonsequentialusers=True, unique_rows=unique_rows)
    np.random.seed(seed)
    ref_batches.split_rows(data, c.val_split, rng=np.random)          # train.py:86 (global stream, before the model)
    twin = init_model()
    rng = np.random.RandomState()
    rng.set_state(np.random.get_state())
    ref = ref_model.RefModel(c.numlayers, c.num_hidden_units, len(unique_cols), c.batch_size,
                             dense_activation=c.activation_type, use_causal_info=c.use_causal_info,
                             use_both_masks=c.auxilliary_mask_type == "both", dropout_probability=c.dropout_probability,
                             dtype=np.float32, rng=np.random.RandomState(0))
    ref.set_weights(twin.model.get_weights())
    ref.dropout_seed = twin.dropout_seed
    ref.compile(ref_model.RefOptimizer("adagrad", lr=c.learning_rate), c.model_loss, rating_range=rating_range)
    B = c.batch_size

    def gen(which, sparsity, **kw):
        return ref_batches.batch_stream(data, B, sparsity, which, c.shuffle_data_every_epoch, c.auxilliary_mask_type,
                                        c.aux_var_value, rng=rng, vectorised=True, **kw)

    history, min_loss, best, best_weights = [], None, 0, None
    for _ in range(c.max_epochs):
        h = ref.fit_generator(gen("train", c.train_sparsity, pass_through_input_training=c.pass_through_input_training),
                              np.floor(data.train_set_size / B) - 1, validation_data=gen("valid", c.train_sparsity),
                              validation_steps=np.floor(data.val_set_size / B) - 1)
        history.append({k: v[-1] for k, v in h.items()})
        val = history[-1][c.early_stopping_metric]
        if not best_weights or val < min_loss:
            min_loss, best, best_weights = val, len(history) - 1, ref.get_weights()
    ref.set_weights(best_weights)
    test = {}
    for s in c.test_sparsities:                                         # train.py:204-213
        vals = ref.evaluate_generator(gen("test", [s, s]), np.floor(data.test_set_size / B) - 1)
        test[s] = dict(zip(ref_model.METRIC_NAMES, vals))
    return {"history": history, "best_epoch": best, "test": test}


def oracle_train_run(fs, cfg, seed, init_model, data=None, n_cols=None, rating_range=None):
    """`train.py:147-177,215-254` driven through the oracle alone on one RandomState stream: the
    per-epoch histories, the fixed-split test metrics and the manual test RMSE a `train.run` of the
    same config must reproduce. `init_model()` builds the product model after `np.random.seed(seed)`
    (its weights, dropout seed and the stream position it leaves are taken over)."""
    from oracle import ref_model
    from omnidirectional_collaborative_filtering_b200 import synthetic
    c = cfg
    if data is None:                               # a synthetic FixedSplit; else a ready RefData (e.g. loaded from files)
        dicts = synthetic.to_reference_dicts(fs, raw_col_id=lambda col: col)
        data = ref_batches.RefData(fs.n_cols, fs.train.n_rows, dicts["unique_cols"], eval_mode="fixed_split",
                                   train=dicts["train"], valid=tuple(dicts["valid"]), test=tuple(dicts["test"]))
        n_cols, rating_range = fs.n_cols, fs.rating_range
    np.random.seed(seed)
    twin = init_model()
    rng = np.random.RandomState()
    rng.set_state(np.random.get_state())          # the stream right after the model's initialisation
    ref = ref_model.RefModel(c.numlayers, c.num_hidden_units, n_cols, c.batch_size,
                             dense_activation=c.activation_type, use_causal_info=c.use_causal_info,
                             use_both_masks=c.auxilliary_mask_type == "both",
                             l2_weight_regulatization=c.l2_weight_regulatization,
                             dropout_probability=c.dropout_probability, dtype=np.float32,
                             rng=np.random.RandomState(0))     # its own draws must not touch either stream
    ref.set_weights(twin.model.get_weights())
    ref.dropout_seed = twin.dropout_seed
    ref.trainable = list(twin.trainable)          # frozen layers of a nested denoising AE (model.py:158-170)
    ref.compile(ref_model.RefOptimizer("adagrad", lr=c.learning_rate), c.model_loss, rating_range=rating_range)
    B = c.batch_size

    def gen(which, sparsity, **kw):
        return ref_batches.batch_stream(data, B, sparsity, which, c.shuffle_data_every_epoch, c.auxilliary_mask_type,
                                        c.aux_var_value, rng=rng, vectorised=True, **kw)

    history, min_loss, best, best_weights = [], None, 0, None
    for _ in range(c.max_epochs):          # callers pick max_epochs/patience so that no early stop happens
        h = ref.fit_generator(gen("train", c.train_sparsity, pass_through_input_training=c.pass_through_input_training),
                              np.floor(data.train_set_size / B) - 1, validation_data=gen("valid", c.train_sparsity),
                              validation_steps=np.floor(data.val_set_size / B) - 1)
        history.append({k: v[-1] for k, v in h.items()})
        val = history[-1][c.early_stopping_metric]
        if not best_weights or val < min_loss:        # train.py:161-169: strict improvement keeps the weights
            min_loss, best, best_weights = val, len(history) - 1, ref.get_weights()
    ref.set_weights(best_weights)                     # train.py:191: the best-validation model is tested
    test = ref.evaluate_generator(gen("test", None), np.floor(data.test_set_size / B) - 1)
    manual = gen("test", None, return_target_count=True)
    sse, count = 0.0, 0
    for _ in range(int(np.floor(data.test_set_size / B))):
        feed, t, n = next(manual)
        sse += float(np.sum(np.square(np.subtract(ref.predict(feed), t, dtype=np.float64))))
        count += n
    return {"history": history, "best_epoch": best, "test": dict(zip(ref_model.METRIC_NAMES, test)),
            "manual_test_rmse": float(np.sqrt(sse / count)), "weights": ref.get_weights()}


def files_pipeline(tmp_dir, seed=3):
    """CSV -> the product's splitter -> the reference's file layout on disk. Returns (directory, n_items, n_rows,
    RefData over the same files read with json.load, as the reference's reader would)."""
    import json
    import os
    from omnidirectional_collaborative_filtering_b200 import splitter
    csv = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "split", "ml", "ratings.csv")
    np.random.seed(seed)
    d = splitter.split_data(csv, str(tmp_dir) + "/", "movielens", include_timestamps=False, save_users_and_items=True)

    def load(name):
        with open(d + name + ".json") as f:
            return json.load(f)

    items = load("unique_items_list")
    train = load("ratingsByUser_dicts_train")
    data = ref_batches.RefData(len(items), len(train), items, eval_mode="fixed_split", train=train,
                               valid=tuple(load("ratingsByUser_dicts_valid")), test=tuple(load("ratingsByUser_dicts_test")))
    return d, len(items), len(train), data


# ---------------------------------------------------------------------------------------------------
# Full-size workloads: reference-shaped dicts over a FixedSplit WITHOUT one Python list per rating
# ---------------------------------------------------------------------------------------------------
class CsrRows(object):
    """Read-only dict {row index: [[column, rating], ...] or None} over a `synthetic.Csr`, rows built on
    demand (what `json.load` of a ratingsBy*_dicts file gives the reference, `data_reader.py:67-70`; a
    Netflix-sized dict of Python lists would need ~20 GB). Keys are the CSR row numbers in store order."""

    def __init__(self, csr, none=None):
        self.csr, self.none = csr, none

    def keys(self):
        return range(self.csr.n_rows)

    def __len__(self):
        return self.csr.n_rows

    def __iter__(self):
        return iter(range(self.csr.n_rows))

    def __contains__(self, k):
        return 0 <= int(k) < self.csr.n_rows

    def __getitem__(self, k):
        k = int(k)
        if self.none is not None and self.none[k]:
            return None
        c, v = self.csr.row(k)
        return [[int(ci), float(vi)] for ci, vi in zip(c, v)]


def oracle_data_from_split(fs):
    """`ref_batches.RefData` over a FixedSplit with lazily built rows: same keys order, same pairing, same
    RNG consumption as the dict files of the same data (keys are ints here; only their count matters to
    `np.random.permutation`)."""
    return ref_batches.RefData(fs.n_cols, fs.train.n_rows, range(fs.n_cols), eval_mode="fixed_split",
                               train=CsrRows(fs.train),
                               valid=(CsrRows(fs.valid_in, fs.valid_none), CsrRows(fs.valid_tg)),
                               test=(CsrRows(fs.test_in, fs.test_none), CsrRows(fs.test_tg)))


def cached_split(shape, reverse, seed=0):
    """`synthetic.make_fixed_split`, cached on disk under /tmp with bench.py's file names so one box builds
    each full-size data set once for the tests and the bench."""
    import os
    import pickle
    from omnidirectional_collaborative_filtering_b200 import synthetic
    path = "/tmp/ocf_b200_%s_%d_%d.pkl" % (shape, int(reverse), seed)
    if os.path.exists(path):
        try:
            with open(path, "rb") as f:
                return pickle.load(f)
        except Exception:
            pass
    fs = synthetic.make_fixed_split(shape, reverse_user_item_data=reverse, seed=seed)
    try:
        tmp = path + ".%d.tmp" % os.getpid()
        with open(tmp, "wb") as f:
            pickle.dump(fs, f, protocol=4)
        os.replace(tmp, path)
    except Exception:
        pass
    return fs


def mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0
