"""Second opinion on the model oracle (oracle/ref_model.py is 'parity unpinned': Keras/TF are
not available): an independent PyTorch-autograd implementation of the same formulas."""
import numpy as np
import pytest
import torch

from oracle import philox, ref_batches, ref_model
from tests.helpers import oracle_data

ACTS = {"sigmoid": torch.sigmoid, "tanh": torch.tanh, "relu": torch.relu,
        "elu": torch.nn.functional.elu, "selu": torch.selu,
        "softplus": torch.nn.functional.softplus, "linear": lambda z: z}


def _feed(golden_datasets, aux, pass_through, seed=3, B=8):
    data = oracle_data(golden_datasets["rev"], "fixed_split")
    np.random.seed(seed)
    gen = ref_batches.batch_stream(data, B, [0.3, 0.8], "train", True, aux, -1,
                                   pass_through_input_training=pass_through)
    return next(gen), data.num_items


def _torch_loss(weights, feed, targets, model, keeps):
    x0, mask = model._split_feed(feed)
    h = torch.tensor(x0)
    for l in range(model.numlayers):
        h = ACTS[model.activation](h @ weights[2 * l] + weights[2 * l + 1])
        if keeps[l] is not None:
            h = h * torch.tensor(keeps[l].astype(np.float64)) / (1.0 - model.p_drop)
    y = torch.tensor(mask) * (h @ weights[-2] + weights[-1])
    t = torch.tensor(np.asarray(targets, dtype=np.float64))
    if model.loss_kind == "mean_squared_error":
        loss = ((y - t) ** 2).mean(dim=1).mean()
    else:
        loss = (y - t).abs().mean(dim=1).mean()
    if model.l2 is not None:
        loss = loss + model.l2 * sum((weights[2 * l] ** 2).sum() for l in range(model.numlayers + 1))
    return loss, y, t


@pytest.mark.parametrize("act", ["sigmoid", "tanh", "elu", "selu", "softplus", "relu", "linear"])
@pytest.mark.parametrize("aux,layers,l2,loss,pdrop", [
    (None, 1, None, "mean_squared_error", None),
    ("dropout", 2, 0.01, "mean_squared_error", 0.2),
    ("both", 3, None, "mean_absolute_error", 0.5),
    ("causal", 1, 0.1, "mean_squared_error", None),
])
def test_gradients_match_autograd(golden_datasets, act, aux, layers, l2, loss, pdrop):
    (feed, targets), N = _feed(golden_datasets, aux, aux is None)
    m = ref_model.RefModel(layers, 12, N, 8, dense_activation=act, use_causal_info=aux is not None,
                           use_both_masks=aux == "both", l2_weight_regulatization=l2,
                           dropout_probability=pdrop, dtype=np.float64,
                           rng=np.random.RandomState(1))
    for i in range(1, len(m.weights), 2):      # non-zero biases
        m.weights[i] = np.random.RandomState(i).normal(size=m.weights[i].shape) * 0.1
    m.compile("adagrad", loss, rating_range=4.0)
    keeps = m._keep_masks(True, None, 8)
    vals, grads = m.gradients(feed, targets)
    tw = [torch.tensor(w, requires_grad=True) for w in m.weights]
    tl, y, t = _torch_loss(tw, feed, targets, m, keeps)
    tl.backward()
    assert abs(vals[0] - tl.item()) <= 1e-12 * max(1, abs(tl.item()))
    for g, w in zip(grads, tw):
        np.testing.assert_allclose(g, w.grad.numpy(), rtol=1e-9, atol=1e-13)
    # metrics, straight from train.py:102-121
    yv, tv = y.detach().numpy(), t.numpy()
    cnt = np.count_nonzero(tv + yv)
    mse_b = ((yv - tv) ** 2).mean(axis=1)
    mae_b = np.abs(yv - tv).mean(axis=1)
    np.testing.assert_allclose(vals[5], np.mean(mse_b * N * 8 / cnt), rtol=1e-12)
    np.testing.assert_allclose(vals[4], np.mean(np.sqrt(mse_b * N * 8 / cnt)), rtol=1e-12)
    np.testing.assert_allclose(vals[2], np.mean(mae_b * N * 8 / cnt), rtol=1e-12)
    np.testing.assert_allclose(vals[3], vals[2] / 4.0, rtol=1e-12)
    np.testing.assert_allclose(vals[1], mae_b.mean(), rtol=1e-12)


@pytest.mark.parametrize("kind", ["adagrad", "rmsprop", "adam", "sgd"])
def test_optimizers_match_independent_formulas(kind):
    rs = np.random.RandomState(0)
    p0 = rs.normal(size=(5, 7))
    gs = [rs.normal(size=(5, 7)) for _ in range(4)]
    opt = ref_model.RefOptimizer(kind, lr=0.01, epsilon=1e-8, decay=0.1)
    p = p0.copy()
    for g in gs:
        opt.apply([p], [g], [True])
    q = p0.copy()
    a = np.zeros_like(q); m = np.zeros_like(q); v = np.zeros_like(q)
    for it, g in enumerate(gs):
        lr = 0.01 / (1 + 0.1 * it)
        if kind == "sgd":
            q = q - lr * g
        elif kind == "adagrad":
            a = a + g ** 2; q = q - lr * g / (np.sqrt(a) + 1e-8)
        elif kind == "rmsprop":
            a = 0.9 * a + 0.1 * g ** 2; q = q - lr * g / (np.sqrt(a) + 1e-8)
        else:
            t = it + 1
            lr_t = lr * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
            m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g ** 2
            q = q - lr_t * m / (np.sqrt(v) + 1e-8)
    np.testing.assert_allclose(p, q, rtol=1e-12)
    if kind in ("adagrad", "rmsprop"):          # torch.optim implements the same two rules
        tp = torch.tensor(p0.copy(), requires_grad=True)
        topt = (torch.optim.Adagrad([tp], lr=0.01, eps=1e-8) if kind == "adagrad"
                else torch.optim.RMSprop([tp], lr=0.01, alpha=0.9, eps=1e-8))
        opt2 = ref_model.RefOptimizer(kind, lr=0.01, epsilon=1e-8)
        p2 = p0.copy()
        for g in gs:
            tp.grad = torch.tensor(g)
            topt.step()
            opt2.apply([p2], [g], [True])
        np.testing.assert_allclose(p2, tp.detach().numpy(), rtol=1e-10)


def test_frozen_layers_and_transfer():
    donor = ref_model.RefModel(1, 6, 10, 4, use_causal_info=False, rng=np.random.RandomState(0))
    new = ref_model.RefModel(3, 6, 10, 4, use_causal_info=False, rng=np.random.RandomState(1))
    new.load_and_fix_for_denoising_autoencoders(donor)
    assert new.trainable == [False, True, True, False]
    assert np.array_equal(new.weights[0], donor.weights[0])
    assert np.array_equal(new.weights[-2], donor.weights[-2])
    before = new.get_weights()
    x = np.random.RandomState(2).normal(size=(4, 10))
    mask = -np.ones((4, 10))
    new.train_on_batch([x, mask], x)
    after = new.get_weights()
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[-1], after[-1])
    assert not np.array_equal(before[2], after[2])
    new.make_trainable()
    assert new.trainable == [True, True, True, False]


def test_philox_known_answer_and_rate():
    # Random123 known-answer vectors for philox4x32-10
    out = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(o) for o in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    out = philox.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(o) for o in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = philox.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(o) for o in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    keep = philox.dropout_keep(1234, 5, 0, 256, 500, 0.2)
    assert keep.shape == (256, 500) and abs(keep.mean() - 0.8) < 0.01
    assert np.array_equal(keep[7:9], philox.dropout_keep(1234, 5, 0, 2, 500, 0.2, row0=7))
