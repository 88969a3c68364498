"""Model files: what `m.save(...)`, `keras.models.load_model(...)` and `save_weights` exchange in the reference
(`train.py:139,169,181-193`, `model.py:102-107`).

The reference's files are Keras 2.0.4 HDF5 files, written under the exact name it passes (no extension:
`..._epoch_7_bestValidScore`). This module keeps those names and speaks two formats, told apart by the
file's first bytes, never by its name:

  * HDF5 in Keras' layout, when `h5py` is importable (it is NOT part of this image; nothing here needs it):
        /                      attrs: keras_version, backend, model_config (JSON, class_name "Model")
        /model_weights         attrs: layer_names, backend, keras_version
        /model_weights/<layer> attrs: weight_names = [b"<layer>/kernel:0", b"<layer>/bias:0"]
        /model_weights/<layer>/<layer>/kernel:0, .../bias:0        float32 datasets, Keras shapes
    (a weights-only file, `save_weights`, holds the content of /model_weights at its root). Optimizer
    state is neither written nor read: the reference deletes `optimizer_weights` before it reloads a model
    (`train.py:183-189`). Reading takes the Dense layers in `layer_names` order, so a donor saved by the
    reference itself (nested-DAE / fine-tuning flows, `train.py:136-145`) loads; the architecture comes from
    `model_config` when present (units, activation, dropout rate, L2, number of inputs), else from the
    kernel shapes.
  * NumPy `.npz` bytes (arrays arr_0.. in Keras weight order + `config`), the fallback when h5py is absent.

Unverified against a real Keras installation (neither Keras nor h5py can be installed here); the layout
follows Keras 2.0.4's `save_model` / `save_weights_to_hdf5_group` and is exercised through an in-memory
stand-in of the h5py API in `tests/test_checkpoint_format.py`.
"""
from __future__ import annotations

import ast
import json
import os

import numpy as np

HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"
ZIP_MAGIC = b"PK"
KERAS_VERSION = b"2.0.4"
BACKEND = b"tensorflow"


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError:
        return None


def file_format(path):
    """'hdf5' | 'npz' of an existing file, by its first bytes."""
    with open(path, "rb") as f:
        head = f.read(8)
    if head.startswith(HDF5_MAGIC):
        return "hdf5"
    if head.startswith(ZIP_MAGIC):
        return "npz"
    raise ValueError("%s is neither an HDF5 nor an .npz model file" % path)


def resolve(path):
    """The file a model path names: the path itself (the reference's convention), else path + '.npz'
    (files written by earlier versions of this package)."""
    if os.path.exists(path):
        return path
    if os.path.exists(path + ".npz"):
        return path + ".npz"
    raise FileNotFoundError(path)


# ---------------------------------------------------------------------------------------------------
# Keras functional-model description of `omni_model` (model.py:43-99)
# ---------------------------------------------------------------------------------------------------
def dense_layer_names(n_dense):
    return ["dense_%d" % (i + 1) for i in range(n_dense)]


def keras_model_config(cfg):
    """`{'class_name': 'Model', 'config': ...}` of the graph `omni_model.__init__` builds for `cfg`
    (our config dict): InputLayers, Concatenate(s), L x (Dense, Dropout), Dense(N, linear), Multiply."""
    N, B = int(cfg["input_shape"]), int(cfg["batch_size"])
    widths = cfg["num_hidden_units"] if isinstance(cfg["num_hidden_units"], (list, tuple)) else [cfg["num_hidden_units"]] * cfg["numlayers"]
    act, l2, p = cfg["dense_activation"], cfg["l2_weight_regulatization"], cfg["dropout_probability"]
    layers = []

    def add(cls, name, conf, inbound):
        layers.append({"class_name": cls, "name": name, "config": dict(conf, name=name),
                       "inbound_nodes": [[[src, 0, 0, {}] for src in inbound]] if inbound else []})
        return name

    def inp(k):
        return add("InputLayer", "input_%d" % k, {"batch_input_shape": [None, N], "dtype": "float32", "sparse": False}, None)

    data, mask = inp(1), inp(2)                       # dataVars, output_mask (model.py:43-45)
    inputs, x, k = [data], data, 3
    if cfg["use_causal_info"]:
        observed = inp(k); k += 1
        x = add("Concatenate", "concatenate_1", {"axis": -1, "trainable": True}, [x, observed])
        inputs.append(observed)
    inputs.append(mask)
    if cfg["use_both_masks"]:
        second = inp(k); k += 1
        x = add("Concatenate", "concatenate_%d" % (2 if cfg["use_causal_info"] else 1), {"axis": -1, "trainable": True}, [x, second])
        inputs.append(second)

    def dense(i, units, activation):
        reg = None if l2 is None else {"class_name": "L1L2", "config": {"l1": 0.0, "l2": float(l2)}}
        return {"units": int(units), "activation": activation or "linear", "use_bias": True, "trainable": True,
                "kernel_initializer": {"class_name": "VarianceScaling",
                                       "config": {"scale": 1.0, "mode": "fan_avg", "distribution": "uniform", "seed": None}},
                "bias_initializer": {"class_name": "Zeros", "config": {}}, "kernel_regularizer": reg,
                "bias_regularizer": None, "activity_regularizer": None, "kernel_constraint": None, "bias_constraint": None}

    for i, wd in enumerate(widths):
        x = add("Dense", "dense_%d" % (i + 1), dense(i, wd, act), [x])
        if p is not None:
            x = add("Dropout", "dropout_%d" % (i + 1), {"rate": float(p), "noise_shape": [B, int(wd)], "trainable": True}, [x])
    full = add("Dense", "dense_%d" % (len(widths) + 1), dense(len(widths), N, "linear"), [x])
    out = add("Multiply", "multiply_1", {"trainable": True}, [mask, full])
    return {"class_name": "Model",
            "config": {"name": "model_1", "layers": layers, "input_layers": [[n, 0, 0] for n in inputs],
                       "output_layers": [[out, 0, 0]]}}


def config_from_keras(model_config, kernels):
    """Our config dict from a Keras `model_config` (may be None) and the Dense kernels in order."""
    N = int(kernels[-1].shape[1])
    widths = [int(k.shape[1]) for k in kernels[:-1]]
    k_blocks = int(kernels[0].shape[0]) // N
    cfg = dict(numlayers=len(widths), num_hidden_units=widths, input_shape=N, batch_size=128, dense_activation="tanh",
               use_causal_info=k_blocks >= 2, use_both_masks=k_blocks >= 3, l2_weight_regulatization=None,
               dropout_probability=None, auxilliary_mask_type="default")
    if model_config is not None:
        layers = model_config.get("config", {}).get("layers", [])
        dense = [l for l in layers if l.get("class_name") == "Dense"]
        drops = [l for l in layers if l.get("class_name") == "Dropout"]
        if dense:
            c = dense[0]["config"]
            cfg["dense_activation"] = c.get("activation", "tanh")
            reg = c.get("kernel_regularizer") or c.get("W_regularizer")
            if reg:
                cfg["l2_weight_regulatization"] = float(reg.get("config", reg).get("l2", 0.0)) or None
        if drops:
            c = drops[0]["config"]
            cfg["dropout_probability"] = float(c.get("rate", c.get("p", 0.0)))
            ns = c.get("noise_shape")
            if ns:
                cfg["batch_size"] = int(ns[0])
    return cfg


# ---------------------------------------------------------------------------------------------------
# writers
# ---------------------------------------------------------------------------------------------------
def _write_weights_group(group, weights):
    names = dense_layer_names(len(weights) // 2)
    group.attrs["layer_names"] = np.array([n.encode("utf8") for n in names])
    group.attrs["backend"] = BACKEND
    group.attrs["keras_version"] = KERAS_VERSION
    for i, name in enumerate(names):
        g = group.create_group(name)
        wnames = ["%s/kernel:0" % name, "%s/bias:0" % name]
        g.attrs["weight_names"] = np.array([w.encode("utf8") for w in wnames])
        for wn, arr in zip(wnames, (weights[2 * i], weights[2 * i + 1])):
            arr = np.ascontiguousarray(arr, dtype=np.float32)
            d = g.create_dataset(wn, arr.shape, dtype=arr.dtype)
            d[...] = arr


def save(path, weights, config=None, fmt=None):
    """Write a model (config given: `m.save`) or its weights only (`save_weights`) to exactly `path`.
    fmt: None = HDF5 when h5py is importable, else npz; or force 'hdf5' / 'npz'."""
    h5 = _h5py()
    if fmt is None:
        fmt = "hdf5" if h5 is not None else "npz"
    if fmt == "hdf5":
        if h5 is None:
            raise RuntimeError("h5py is not installed: cannot write an HDF5 model file")
        with h5.File(path, "w") as f:
            if config is None:
                _write_weights_group(f, weights)
            else:
                f.attrs["keras_version"] = KERAS_VERSION
                f.attrs["backend"] = BACKEND
                f.attrs["model_config"] = json.dumps(keras_model_config(config)).encode("utf8")
                f.attrs["ocf_config"] = repr(config).encode("utf8")       # ours: what Keras' config cannot carry (mask type, width list)
                _write_weights_group(f.create_group("model_weights"), weights)
        return path
    if fmt != "npz":
        raise ValueError("fmt must be 'hdf5' or 'npz'")
    with open(path, "wb") as f:                     # a file object: np.savez must not append '.npz' to the reference's name
        if config is None:
            np.savez(f, *weights)
        else:
            np.savez(f, *weights, config=np.array(repr(config)))
    return path


# ---------------------------------------------------------------------------------------------------
# readers
# ---------------------------------------------------------------------------------------------------
def _text(v):
    if isinstance(v, np.ndarray) and v.shape == ():
        v = v.item()
    return v.decode("utf8") if isinstance(v, (bytes, np.bytes_)) else str(v)


def _read_hdf5(path):
    h5 = _h5py()
    if h5 is None:
        raise RuntimeError("%s is an HDF5 (Keras) model file; reading it needs h5py, which is not installed" % path)
    with h5.File(path, "r") as f:
        is_model = "model_weights" in f                 # else a weights-only file (`save_weights`)
        group = f["model_weights"] if is_model else f
        model_config = json.loads(_text(f.attrs["model_config"])) if "model_config" in f.attrs else None
        own = ast.literal_eval(_text(f.attrs["ocf_config"])) if "ocf_config" in f.attrs else None
        weights = []
        for name in [_text(n) for n in group.attrs["layer_names"]]:
            g = group[name]
            wnames = [_text(w) for w in g.attrs["weight_names"]] if "weight_names" in g.attrs else []
            arrs = [np.asarray(g[w], dtype=np.float32) for w in wnames]
            if len(arrs) == 2 and arrs[0].ndim == 2 and arrs[1].ndim == 1:        # a Dense layer: kernel, bias
                weights += arrs
            elif arrs:
                raise ValueError("layer %r of %s holds weights this architecture does not have" % (name, path))
    if not weights:
        raise ValueError("%s holds no Dense layers" % path)
    if own is None and is_model:
        own = config_from_keras(model_config, weights[0::2])
    return own, weights


def _read_npz(path):
    with np.load(path, allow_pickle=False) as f:
        cfg = ast.literal_eval(str(f["config"])) if "config" in f.files else None
        keys = sorted((k for k in f.files if k.startswith("arr_")), key=lambda k: int(k[4:]))
        weights = [f[k] for k in keys]
    return cfg, weights


def load(path):
    """(config dict or None, weights in Keras order) of a model or weights file in either format."""
    path = resolve(path)
    return _read_hdf5(path) if file_format(path) == "hdf5" else _read_npz(path)
