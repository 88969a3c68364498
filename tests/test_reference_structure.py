"""Model structure and weight transfer / freezing (row a13) against the REFERENCE's own code:
`tests/golden/transfer.json` was produced by `/root/reference/model.py` itself (constructor + the four
transfer helpers, model.py:33-170) executed over a structural Keras stand-in
(`tests/golden/make_transfer_golden.py`). The product's `omni_model` (host weights, no GPU) and the oracle's
`RefModel` must place the same donor layers in the same slots with the same `trainable` flags, and expose the
same layer shapes and input order."""
import json
import os

import numpy as np
import pytest

from oracle import ref_model
from omnidirectional_collaborative_filtering_b200.model import omni_model
from tests.conftest import GOLDEN

with open(os.path.join(GOLDEN, "transfer.json")) as _f:
    CASES = json.load(_f)


def _tag(weights, model_id):
    out = []
    for i, w in enumerate(weights):
        out.append(np.full(w.shape, 100 * model_id + i // 2 + (0.5 if i % 2 else 0.0), dtype=np.float32))
    return out


def _read(weights, trainable):
    out = []
    for l in range(len(weights) // 2):
        k, b = weights[2 * l], weights[2 * l + 1]
        assert np.all(k == k.flat[0]) and np.all(b == b.flat[0]) and b.flat[0] == k.flat[0] + 0.5
        out.append({"from_model": int(k.flat[0]) // 100, "from_dense": int(k.flat[0]) % 100, "trainable": bool(trainable[l])})
    return out


class _Product(object):
    def __init__(self, L, H, N, **kw):
        self.om = omni_model(L, H, N, 16, dense_activation="sigmoid", **kw)

    def tag(self, model_id):
        self.om.model.set_weights(_tag(self.om.model.get_weights(), model_id))

    def read(self):
        return _read(self.om.model.get_weights(), self.om.trainable)

    donor = property(lambda self: self.om.model)
    target = property(lambda self: self.om)


class _Oracle(object):
    def __init__(self, L, H, N, use_causal_info=True, dropout_probability=None, **kw):
        self.om = ref_model.RefModel(L, H, N, 16, dense_activation="sigmoid", use_causal_info=use_causal_info,
                                     dropout_probability=dropout_probability, dtype=np.float32, rng=np.random.RandomState(0))

    def tag(self, model_id):
        self.om.set_weights(_tag(self.om.get_weights(), model_id))

    def read(self):
        return _read(self.om.get_weights(), self.om.trainable)

    donor = property(lambda self: self.om)
    target = property(lambda self: self.om)


TRANSFER = [c for c in CASES if c["kind"] != "structure"]


@pytest.mark.parametrize("impl", [_Product, _Oracle], ids=["product", "oracle"])
def test_weight_transfer_places_layers_like_the_reference(impl, capsys):
    assert len(TRANSFER) >= 80
    for c in TRANSFER:
        H, N = c["H"], c["N"]
        if c["kind"].startswith("load_and_fix"):
            kw = dict(use_causal_info=False, dropout_probability=c["dropout"])
            donor, new = impl(c["donor_layers"], H, N, **kw), impl(c["new_layers"], H, N, **kw)
            donor.tag(1), new.tag(2)
            new.target.load_and_fix_for_denoising_autoencoders(donor.donor)
            if c["kind"].endswith("make_trainable"):
                new.target.make_trainable()
        elif c["kind"] == "manually_load_all":
            donor, new = impl(c["layers"], H, N, use_causal_info=True), impl(c["layers"], H, N, use_causal_info=True)
            donor.tag(1), new.tag(2)
            new.target.manually_load_all_weights(donor.donor)
        else:
            donor, new = impl(c["layers"], H, N, use_causal_info=False), impl(c["layers"], H, N, use_causal_info=False)
            donor.tag(1), new.tag(2)
            new.target.replace_dense_layer_weights(donor.donor, c["mask"], make_layers_trainable=c["make_trainable"])
        assert new.read() == c["result"], c
    capsys.readouterr()


@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "structure"],
                         ids=lambda c: "L%d-%s" % (c["numlayers"], "-".join(sorted(k for k, v in c["kwargs"].items() if v))))
def test_graph_structure_matches_the_reference(case):
    kw, res = case["kwargs"], case["result"]
    L, H, N = case["numlayers"], case["H"], case["N"]
    om = omni_model(L, H, N, 16, dense_activation="sigmoid", **kw)
    dense = [(tuple(i), tuple(o)) for name, i, o in res["layers"] if name.startswith("dense")]
    # the Dense chain: [k*N -> H], (L-1) x [H -> H], [H -> N]; k = number of concatenated input blocks
    assert [(i[1], o[1]) for i, o in dense] == om.weight_shapes()[0::2]
    assert [o[1] for _, o in dense] == [s[0] for s in om.weight_shapes()[1::2]]
    ref = ref_model.RefModel(L, H, N, 16, dense_activation="sigmoid", dtype=np.float32, rng=np.random.RandomState(0),
                             use_causal_info=kw.get("use_causal_info", True), use_both_masks=kw.get("use_both_masks", False),
                             l2_weight_regulatization=kw.get("l2_weight_regulatization"),
                             dropout_probability=kw.get("dropout_probability"))
    assert [w.shape for w in ref.get_weights()] == [tuple(s) for s in om.weight_shapes()]
    # inputs in the order the generator feeds them (data_reader.py:354-361): data, (aux), output mask, (second mask);
    # input_1 = data, input_2 = output mask, input_3 = aux, input_4 = second mask (creation order, model.py:43-56)
    want = ["input_1"] + (["input_3"] if kw.get("use_causal_info") else []) + ["input_2"] + (["input_4"] if kw.get("use_both_masks") else [])
    assert res["inputs"] == want
    concat = [i for name, i, o in res["layers"] if name.startswith("concatenate")]
    assert len(concat) == om.k_blocks - 1              # x0 = [data | aux | second]: k blocks of W_enc rows
    # one Dropout(p, noise_shape=[B, H]) behind every hidden layer, none behind the output layer (model.py:72-73)
    assert res["dropout_noise_shapes"] == ([[16, H]] * L if kw.get("dropout_probability") is not None else [])
    # the L2 regulariser sits on EVERY Dense kernel, the output layer's included, never on a bias (model.py:65-66,81-82)
    lam = kw.get("l2_weight_regulatization")
    assert res["regularized"] == [lam] * (L + 1)
    if lam is not None:
        expect = lam * sum(float(np.sum(np.square(w.astype(np.float64)))) for w in ref.get_weights()[0::2])
        assert float(ref._reg()) == pytest.approx(expect, rel=1e-5)


def test_trainable_flags_after_compile_both_keras_behaviours():
    """model.py:109-170 set layer.trainable AFTER m.compile (train.py:131-145). Default: the flags count at once (the
    flows' intent). trainable_applies = "on_compile": Keras 2.0.4 collected the trainable weights at compile time, so the
    flags wait for the next compile (model.py:137's own remark)."""
    from omnidirectional_collaborative_filtering_b200.model import omni_model
    np.random.seed(3)
    donor = omni_model(1, 8, 12, 4, use_causal_info=False)
    for mode, after_helper, after_compile in (("at_once", [False, True, False], [False, True, False]),
                                              ("on_compile", [True, True, True], [False, True, False])):
        om = omni_model(2, 8, 12, 4, use_causal_info=False)
        om.trainable_applies = mode
        om.model.compile("adagrad", "mean_squared_error")
        om.load_and_fix_for_denoising_autoencoders(donor)
        assert om.trainable == [False, True, False]                   # layer.trainable as the helper left it
        assert om._compiled_trainable == after_helper
        om.model.compile("adagrad", "mean_squared_error")
        assert om._compiled_trainable == after_compile
