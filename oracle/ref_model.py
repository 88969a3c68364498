"""ORACLE (test infrastructure, not product code): dense NumPy transcription of the reference's
model, loss, metrics, optimizers and epoch loop.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.

Parity status: PARITY UNPINNED for the third-party arithmetic, pinned for everything the reference
itself defines. The layer arithmetic, autodiff and optimizer rules of this path live in Keras 2.0.4
on TensorFlow 1.3.0 (`/root/reference/README.md:17-23`), neither of which is in `/root/reference`
nor installable offline, and the reference has no tests, golden vectors or logged results: Dense /
activation / Dropout forward and backward and the Adagrad / RMSprop / Adam / SGD rules are restated
from those releases' published algorithms, and `tests/test_oracle_model.py` cross-checks them against
an independent PyTorch-autograd implementation of the same formulas. What the reference's OWN code
defines is pinned by running that code here (fixtures under `tests/golden/`, each with the script
that made it):
  * the four custom metrics          train.py:102-121, closures cut out with `ast`     tests/test_reference_metrics.py
  * graph structure, input order,    model.py:33-99 executed over a structural          tests/test_reference_structure.py
    dropout placement, L2 placement  Keras stand-in
  * the weight-transfer helpers      model.py:109-170, same run (88 cases)              tests/test_reference_structure.py
  * epoch loop, early stopping,      train.py:147-255 executed over recording           tests/test_train_host.py
    test procedures, manual RMSE     stand-ins (14 call traces)

What is restated (reference file:line):
  RefModel.__init__/forward   graph of `omni_model`, `model.py:43-99`:
                              x = concat(data,(aux),(second)) :47-60; L x Dense(H, act) +
                              Dropout(p, [B,H]) :64-73; Dense(N, linear) :82-84; y = mask*full :86
  loss                        `train.py:49,131-132` mean_squared_error / mean_absolute_error
                              = mean over N then mean over B, plus l2 * sum(W^2) per kernel
                              (`model.py:65-66,81-82`)
  metrics                     `train.py:102-121` accurate_MAE/RMSE/MSE, nMAE, and 'mae'
  optimizers                  Keras 2.0.4 Adagrad / RMSprop / Adam / SGD update rules at
                              `train.py:50-51`, `train_jester.py:61`
  fit_generator / evaluate_generator / predict   call sites `train.py:157-158,208,218,239`:
                              per-epoch value = arithmetic mean of per-batch values; training
                              values are computed before the update with dropout active;
                              validation has dropout off
  run_training                epoch loop + early stopping `train.py:147-177`; tests `:202-254`
  transfer helpers            `model.py:109-170`
"""
from __future__ import annotations

import numpy as np

from . import philox

SELU_ALPHA = 1.6732632423543772
SELU_SCALE = 1.0507009873554805


def act_forward(name, z):
    if name == "linear" or name is None:
        return z
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    if name == "tanh":
        return np.tanh(z)
    if name == "relu":
        return np.maximum(z, 0)
    if name == "elu":
        return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))
    if name == "selu":
        return SELU_SCALE * np.where(z > 0, z, SELU_ALPHA * np.expm1(np.minimum(z, 0)))
    if name == "softplus":
        return np.logaddexp(z, 0)
    raise ValueError("activation %r" % (name,))


def act_backward(name, z, a):
    """d a / d z given pre-activation z and activation a."""
    if name == "linear" or name is None:
        return np.ones_like(z)
    if name == "sigmoid":
        return a * (1 - a)
    if name == "tanh":
        return 1 - a * a
    if name == "relu":
        return (z > 0).astype(z.dtype)
    if name == "elu":
        return np.where(z > 0, 1.0, np.exp(np.minimum(z, 0))).astype(z.dtype)
    if name == "selu":
        return (SELU_SCALE * np.where(z > 0, 1.0, SELU_ALPHA * np.exp(np.minimum(z, 0)))).astype(z.dtype)
    if name == "softplus":
        return (1.0 / (1.0 + np.exp(-z))).astype(z.dtype)
    raise ValueError("activation %r" % (name,))


class RefOptimizer(object):
    """Keras 2.0.4 update rules."""

    def __init__(self, kind="adagrad", lr=None, epsilon=1e-8, decay=0.0, rho=0.9,
                 beta_1=0.9, beta_2=0.999):
        kind = kind.lower()
        defaults = {"adagrad": 0.01, "rmsprop": 0.001, "adam": 0.001, "sgd": 0.01}
        self.kind = kind
        self.lr = defaults[kind] if lr is None else lr
        self.epsilon, self.decay, self.rho = epsilon, decay, rho
        self.beta_1, self.beta_2 = beta_1, beta_2
        self.iterations = 0
        self.state = {}

    def reset(self):
        self.iterations = 0
        self.state = {}

    def apply(self, params, grads, trainable):
        """In-place update of the arrays in `params` (list), `grads` aligned, `trainable` bools."""
        dt = params[0].dtype.type
        lr = self.lr
        if self.decay > 0:
            lr = lr * (1.0 / (1.0 + self.decay * self.iterations))
        t = self.iterations + 1
        eps = dt(self.epsilon)
        for idx, (p, g, tr) in enumerate(zip(params, grads, trainable)):
            if not tr:
                continue
            if self.kind == "sgd":
                p -= dt(lr) * g
            elif self.kind == "adagrad":
                a = self.state.setdefault(idx, np.zeros_like(p))
                a += g * g
                p -= dt(lr) * g / (np.sqrt(a) + eps)
            elif self.kind == "rmsprop":
                a = self.state.setdefault(idx, np.zeros_like(p))
                a *= dt(self.rho)
                a += dt(1.0 - self.rho) * g * g
                p -= dt(lr) * g / (np.sqrt(a) + eps)
            elif self.kind == "adam":
                m, v = self.state.setdefault(idx, (np.zeros_like(p), np.zeros_like(p)))
                lr_t = lr * (np.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t))
                m *= dt(self.beta_1)
                m += dt(1.0 - self.beta_1) * g
                v *= dt(self.beta_2)
                v += dt(1.0 - self.beta_2) * g * g
                p -= dt(lr_t) * m / (np.sqrt(v) + eps)
            else:
                raise ValueError(self.kind)
        self.iterations += 1


METRIC_NAMES = ["loss", "mean_absolute_error", "accurate_MAE", "nMAE", "accurate_RMSE", "accurate_MSE"]


class RefModel(object):
    """`omni_model` + the Keras `Model` methods the reference calls on it."""

    def __init__(self, numlayers, num_hidden_units, input_shape, batch_size,
                 dense_activation="tanh", use_causal_info=True, use_timestamps=False,
                 use_both_masks=False, l2_weight_regulatization=None,
                 sparse_representation=False, dropout_probability=None,
                 use_sparse_masking_layer=False, dtype=np.float32, rng=np.random):
        if use_timestamps or sparse_representation or use_sparse_masking_layer:
            raise NotImplementedError("out of scope (SURVEY.md section 2, rows 11-13)")
        self.numlayers = int(numlayers)
        if isinstance(num_hidden_units, (list, tuple)):
            self.widths = [int(w) for w in num_hidden_units]
            assert len(self.widths) == self.numlayers
        else:
            self.widths = [int(num_hidden_units)] * self.numlayers
        self.num_hidden_units = self.widths[0] if self.widths else 0
        self.input_shape = int(input_shape)
        self.batch_size = int(batch_size)
        self.activation = dense_activation
        self.k_blocks = 1 + int(bool(use_causal_info)) + int(bool(use_both_masks))
        self.use_causal_info, self.use_both_masks = bool(use_causal_info), bool(use_both_masks)
        self.l2 = l2_weight_regulatization
        self.p_drop = dropout_probability
        self.dtype = np.dtype(dtype)
        dims = [self.k_blocks * self.input_shape] + self.widths + [self.input_shape]
        self.weights = []
        for fan_in, fan_out in zip(dims[:-1], dims[1:]):
            lim = np.sqrt(6.0 / (fan_in + fan_out))              # glorot_uniform [3P]
            self.weights.append(rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(self.dtype))
            self.weights.append(np.zeros(fan_out, dtype=self.dtype))
        self.trainable = [True] * (self.numlayers + 1)           # per dense layer
        self.optimizer = RefOptimizer("adagrad", 0.005)
        self.loss_kind = "mean_squared_error"
        self.rating_range = 1.0
        self.dropout_seed = 0
        self.step_counter = 0
        self.model = self                                        # `omni_m.model` in train.py:99

    # -- Keras Model surface -------------------------------------------------------------
    def compile(self, optimizer="adagrad", loss="mean_squared_error", metrics=None,
                rating_range=1.0):
        self.optimizer = RefOptimizer(optimizer) if isinstance(optimizer, str) else optimizer
        self.loss_kind = loss
        self.rating_range = rating_range

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, weights):
        assert len(weights) == len(self.weights)
        for i, w in enumerate(weights):
            assert w.shape == self.weights[i].shape
            self.weights[i] = np.array(w, dtype=self.dtype)

    load_weights = set_weights                                    # model.py:106-107

    def _split_feed(self, feed):
        """feed = [x, (aux), mask_out, (second)] (data_reader.py:354-361) -> (x0, mask_out)."""
        if self.use_causal_info:
            parts = [feed[0], feed[1]]
            mask = feed[2]
            if self.use_both_masks:
                parts.append(feed[3])
        else:
            parts = [feed[0]]
            mask = feed[1]
        x0 = np.concatenate([np.asarray(p, dtype=self.dtype) for p in parts], axis=1)
        return x0, np.asarray(mask, dtype=self.dtype)

    def _keep_masks(self, training, dropout_masks, n_rows):
        if not training or self.p_drop is None:
            return [None] * self.numlayers
        if dropout_masks is not None:
            return list(dropout_masks)
        return [philox.dropout_keep(self.dropout_seed, self.step_counter, l, n_rows,
                                    self.widths[l], self.p_drop) for l in range(self.numlayers)]

    def forward(self, feed, training=False, dropout_masks=None):
        x0, mask = self._split_feed(feed)
        keeps = self._keep_masks(training, dropout_masks, x0.shape[0])
        acts = [x0]
        cache = []
        h = x0
        for l in range(self.numlayers):
            W, b = self.weights[2 * l], self.weights[2 * l + 1]
            z = h @ W + b
            a = act_forward(self.activation, z)
            scale = None
            if keeps[l] is not None:
                scale = keeps[l].astype(self.dtype) / self.dtype.type(1.0 - self.p_drop)
                h = a * scale
            else:
                h = a
            cache.append((z, a, scale))
            acts.append(h)
        full = h @ self.weights[-2] + self.weights[-1]
        y = mask * full
        return y, full, acts, cache, mask

    def _values(self, y, t):
        """[loss-without-reg, mae, accurate_MAE, nMAE, accurate_RMSE, accurate_MSE] (train.py:102-121)."""
        B, N = y.shape
        dt = self.dtype.type
        err = y - t
        sse_b = np.sum(err * err, axis=1, dtype=self.dtype)
        sae_b = np.sum(np.abs(err), axis=1, dtype=self.dtype)
        mse_b, mae_b = sse_b / dt(N), sae_b / dt(N)
        cnt = dt(np.count_nonzero(t + y))
        with np.errstate(divide="ignore", invalid="ignore"):
            acc_mae = np.mean(mae_b * dt(N) * dt(B) / cnt, dtype=self.dtype)
            acc_mse = np.mean(mse_b * dt(N) * dt(B) / cnt, dtype=self.dtype)
            acc_rmse = np.mean(np.sqrt(mse_b * dt(N) * dt(B) / cnt), dtype=self.dtype)
        mae = np.mean(mae_b, dtype=self.dtype)
        loss = np.mean(mse_b, dtype=self.dtype) if self.loss_kind == "mean_squared_error" else mae
        return [loss, mae, acc_mae, acc_mae / dt(self.rating_range), acc_rmse, acc_mse]

    def _reg(self):
        if self.l2 is None:
            return self.dtype.type(0)
        return self.dtype.type(self.l2) * sum(np.sum(self.weights[2 * l] ** 2, dtype=self.dtype)
                                               for l in range(self.numlayers + 1))

    def test_on_batch(self, feed, targets):
        y, _, _, _, _ = self.forward(feed, training=False)
        vals = self._values(y, np.asarray(targets, dtype=self.dtype))
        vals[0] = vals[0] + self._reg()
        return [float(v) for v in vals]

    def predict(self, feed, batch_size=None, verbose=0):
        return self.forward(feed, training=False)[0]

    def score(self, feed):
        """Full-catalogue scores: the tensor before the mask multiply (model.py:82-84)."""
        return self.forward(feed, training=False)[1]

    def gradients(self, feed, targets, dropout_masks=None):
        """(metric values, gradient list aligned with self.weights) of one training batch."""
        t = np.asarray(targets, dtype=self.dtype)
        y, full, acts, cache, mask = self.forward(feed, training=True, dropout_masks=dropout_masks)
        vals = self._values(y, t)
        vals[0] = vals[0] + self._reg()
        B, N = y.shape
        dt = self.dtype.type
        err = y - t
        if self.loss_kind == "mean_squared_error":
            dy = err * dt(2.0 / (B * N))
        else:
            dy = np.sign(err) * dt(1.0 / (B * N))
        dfull = dy * mask
        grads = [None] * len(self.weights)
        h = acts[-1]
        grads[-2] = h.T @ dfull
        grads[-1] = dfull.sum(axis=0)
        dh = dfull @ self.weights[-2].T
        for l in reversed(range(self.numlayers)):
            z, a, scale = cache[l]
            if scale is not None:
                dh = dh * scale
            dz = dh * act_backward(self.activation, z, a)
            grads[2 * l] = acts[l].T @ dz
            grads[2 * l + 1] = dz.sum(axis=0)
            if l > 0:
                dh = dz @ self.weights[2 * l].T
        if self.l2 is not None:
            for l in range(self.numlayers + 1):
                grads[2 * l] = grads[2 * l] + dt(2.0 * self.l2) * self.weights[2 * l]
        return vals, grads

    def train_on_batch(self, feed, targets, dropout_masks=None):
        vals, grads = self.gradients(feed, targets, dropout_masks)
        flags = []
        for l in range(self.numlayers + 1):
            flags += [self.trainable[l], self.trainable[l]]
        self.optimizer.apply(self.weights, grads, flags)
        self.step_counter += 1
        return [float(v) for v in vals]

    def fit_generator(self, generator, steps_per_epoch, validation_data=None, validation_steps=None):
        """One epoch (train.py:157-158). Returns {"history": {name: [value]}}-like dict."""
        steps = int(steps_per_epoch)
        rows = [self.train_on_batch(*next(generator)[:2]) for _ in range(steps)]
        hist = {}
        mean = np.mean(np.asarray(rows, dtype=np.float64), axis=0) if rows else [np.nan] * 6
        for name, v in zip(METRIC_NAMES, mean):
            hist[name] = [float(v)]
        if validation_data is not None:
            vals = self.evaluate_generator(validation_data, validation_steps)
            for name, v in zip(METRIC_NAMES, vals):
                hist["val_" + name] = [float(v)]
        return hist

    def evaluate_generator(self, generator, steps):
        rows = [self.test_on_batch(*next(generator)[:2]) for _ in range(int(steps))]
        return [float(v) for v in np.mean(np.asarray(rows, dtype=np.float64), axis=0)]

    # -- weight transfer (model.py:109-170) ----------------------------------------------
    def dense_layer_weights(self):
        return [[self.weights[2 * l], self.weights[2 * l + 1]] for l in range(self.numlayers + 1)]

    def _set_dense(self, l, pair):
        self.weights[2 * l] = np.array(pair[0], dtype=self.dtype)
        self.weights[2 * l + 1] = np.array(pair[1], dtype=self.dtype)

    def load_and_fix_for_denoising_autoencoders(self, donor):
        """Outer floor(D/2) dense layers on each side come from the donor and freeze (:142-170)."""
        donor_layers = donor.dense_layer_weights()
        n_side = int(len(donor_layers) / 2)
        n_new = self.numlayers + 1
        for l in range(n_new):
            if l < n_side:
                self._set_dense(l, donor_layers[l])
                self.trainable[l] = False
            elif l >= n_new - n_side:
                self._set_dense(l, donor_layers[len(donor_layers) - (n_new - l)])
                self.trainable[l] = False

    def manually_load_all_weights(self, donor):
        self.set_weights(donor.get_weights())                      # :129-134

    def replace_dense_layer_weights(self, donor, layers_to_replace, make_layers_trainable=False):
        donor_layers = donor.dense_layer_weights()                 # :109-127
        if layers_to_replace == "all":
            layers_to_replace = [True] * len(donor_layers)
        for l in range(self.numlayers + 1):
            if layers_to_replace[l]:
                self._set_dense(l, donor_layers[l])
                self.trainable[l] = make_layers_trainable

    def make_trainable(self):
        for l in range(self.numlayers):                            # :136-140 (output width == H)
            self.trainable[l] = True
        if self.input_shape == self.num_hidden_units:
            self.trainable[self.numlayers] = True


def full_rmse(predictions, targets, ratings_count):
    """`compute_full_RMSE`, train.py:243-252."""
    sse = 0.0
    for p, t in zip(predictions, targets):
        sse += np.sum(np.square(np.subtract(p, t)))
    return float(np.sqrt(sse / ratings_count))


def run_training(model, make_gen, train_size, val_size, batch_size, max_epochs, patience,
                 early_stopping_metric="val_accurate_MSE"):
    """Epoch loop with early stopping, train.py:147-177 (generators driven synchronously:
    the train stream of an epoch is consumed before its validation stream is created's first
    batch is drawn; SURVEY.md Appendix B, last row).

    make_gen(which) -> a fresh generator. Returns (history list, best_epoch, best_weights).
    The reference never saves the first epoch as best (`train.py:164-165`); like the product
    we keep its weights so "best" is always defined, the stopping decisions are unchanged.
    """
    min_loss, best_epoch, best_weights = None, 0, None
    history = []
    for i in range(max_epochs):
        train_gen, valid_gen = make_gen("train"), make_gen("valid")
        hist = model.fit_generator(train_gen, np.floor(train_size / batch_size) - 1,
                                   validation_data=valid_gen,
                                   validation_steps=np.floor(val_size / batch_size) - 1)
        history.append(hist)
        val_loss = hist[early_stopping_metric][-1]
        if min_loss is None:
            min_loss, best_weights = val_loss, model.get_weights()
        elif min_loss > val_loss:
            min_loss, best_epoch, best_weights = val_loss, i, model.get_weights()
        elif i - best_epoch > patience:
            break
    return history, best_epoch, best_weights
