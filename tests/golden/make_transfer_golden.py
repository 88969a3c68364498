#!/usr/bin/env python
"""Golden fixture for the model STRUCTURE and the weight-transfer helpers, produced by the reference's own
`model.py` (model.py:33-170) executed over `keras_structure_stub` (Keras / TensorFlow cannot be installed
here; the stub models the layer graph, names, shapes, weights and `trainable` flags, no arithmetic).

    python tests/golden/make_transfer_golden.py        # needs /root/reference; writes tests/golden/transfer.json

Every Dense layer's kernel and bias are filled with a tag (model id, dense index) before a helper runs; afterwards
the tags say which donor layer each layer of the recipient holds, next to its `trainable` flag."""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import keras_structure_stub as stub  # noqa: E402


def load_reference_model():
    stub.install()
    spec = importlib.util.spec_from_file_location("ref_model_py", "/root/reference/model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def tag(om, model_id):
    dense = [l for l in om.model.layers if isinstance(l, stub.Dense)]
    for j, l in enumerate(dense):
        l.kernel[:] = 100 * model_id + j
        l.bias[:] = 100 * model_id + j + 0.5
    return dense


def read(dense):
    out = []
    for l in dense:
        k, b = float(l.kernel.flat[0]), float(l.bias.flat[0])
        assert np.all(l.kernel == k) and np.all(l.bias == b) and b == k + 0.5
        out.append({"from_model": int(k) // 100, "from_dense": int(k) % 100, "trainable": bool(l.trainable)})
    return out


def structure(om):
    m = om.model
    return {"layers": [[l.name, l.input_shape, l.output_shape] for l in m.layers],
            "inputs": [t.layer.name for t in m.inputs],
            "dropout_noise_shapes": [l.noise_shape for l in m.layers if isinstance(l, stub.Dropout)],
            "regularized": [l.kernel_regularizer.value if l.kernel_regularizer else None
                            for l in m.layers if isinstance(l, stub.Dense)]}


def main():
    ref = load_reference_model()
    cases = []

    def build(L, H, N, **kw):
        stub.reset_names()
        return ref.omni_model(L, H, N, 16, **kw)

    # structure of the graph (a6): inputs, concat order, dense / dropout chain, regularisers
    for kw in (dict(use_causal_info=False), dict(use_causal_info=True), dict(use_causal_info=True, use_both_masks=True),
               dict(use_causal_info=False, dropout_probability=0.2, l2_weight_regulatization=0.01)):
        for L in (1, 3):
            om = build(L, 8, 20, dense_activation="sigmoid", **kw)
            cases.append({"kind": "structure", "numlayers": L, "H": 8, "N": 20, "kwargs": kw, "result": structure(om)})

    # weight transfer (a13)
    for H, N in ((8, 20), (8, 8)):
        for Ld, Ln in ((1, 3), (1, 2), (2, 4), (3, 3), (1, 1), (2, 3), (3, 5), (2, 2)):
            for drop in (None, 0.2):
                kw = dict(use_causal_info=False, dropout_probability=drop)
                donor, new = build(Ld, H, N, **kw), build(Ln, H, N, **kw)
                tag(donor, 1)
                dense = tag(new, 2)
                try:
                    new.load_and_fix_for_denoising_autoencoders(donor.model)
                    res = read(dense)
                except ValueError as e:                    # shape mismatch inside set_weights
                    res = "ValueError"
                cases.append({"kind": "load_and_fix", "H": H, "N": N, "donor_layers": Ld, "new_layers": Ln,
                              "dropout": drop, "result": res})
                if res != "ValueError":
                    new.make_trainable()
                    cases.append({"kind": "load_and_fix+make_trainable", "H": H, "N": N, "donor_layers": Ld,
                                  "new_layers": Ln, "dropout": drop, "result": read(dense)})
        for L in (1, 2, 3):
            donor, new = build(L, H, N, use_causal_info=True), build(L, H, N, use_causal_info=True)
            tag(donor, 1)
            dense = tag(new, 2)
            new.manually_load_all_weights(donor.model)
            cases.append({"kind": "manually_load_all", "H": H, "N": N, "layers": L, "result": read(dense)})
            for mask, trainable in (("all", False), ([True] + [False] * L, True), ([False] * L + [True], False)):
                donor, new = build(L, H, N, use_causal_info=False), build(L, H, N, use_causal_info=False)
                tag(donor, 1)
                dense = tag(new, 2)
                new.replace_dense_layer_weights(donor.model, mask, make_layers_trainable=trainable)
                cases.append({"kind": "replace", "H": H, "N": N, "layers": L, "mask": mask, "make_trainable": trainable,
                              "result": read(dense)})
    with open(os.path.join(HERE, "transfer.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
