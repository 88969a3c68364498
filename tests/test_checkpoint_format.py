"""Model-file formats (`checkpoint.py`; train.py:139,169,181-193): exact file names, format sniffing, the Keras HDF5
layout through a stand-in of the h5py API (h5py itself is not in this image), a reference-style donor with
non-Dense layers in `layer_names`, and the npz fallback."""
import json
import sys

import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import checkpoint
from omnidirectional_collaborative_filtering_b200.model import load_model, omni_model
from tests import h5py_stub


@pytest.fixture
def fake_h5py(monkeypatch):
    monkeypatch.setitem(sys.modules, "h5py", h5py_stub)
    return h5py_stub


def _model(**kw):
    np.random.seed(3)
    base = dict(dense_activation="elu", use_causal_info=True, use_both_masks=True, l2_weight_regulatization=0.01,
                dropout_probability=0.3, auxilliary_mask_type="both")
    base.update(kw)
    return omni_model(2, [12, 6], 30, 8, **base)


def test_npz_fallback_keeps_the_references_file_name(tmp_path, monkeypatch):
    monkeypatch.setitem(sys.modules, "h5py", None)              # import h5py -> ImportError
    om = _model()
    path = str(tmp_path / "stackedDenoising_epoch_3_bestValidScore")        # the reference's names carry no extension
    om.model.save(path)
    assert (tmp_path / "stackedDenoising_epoch_3_bestValidScore").exists() and not (tmp_path / "stackedDenoising_epoch_3_bestValidScore.npz").exists()
    assert checkpoint.file_format(path) == "npz"
    back = load_model(path)
    assert back.owner.config() == om.config()
    for a, b in zip(back.get_weights(), om.model.get_weights()):
        assert np.array_equal(a, b)
    # files of earlier versions (name + ".npz") still load through the bare name
    np.savez(str(tmp_path / "old.npz"), *om.model.get_weights(), config=np.array(repr(om.config())))
    assert load_model(str(tmp_path / "old")).owner.config() == om.config()


def test_hdf5_layout_is_keras_and_round_trips(tmp_path, fake_h5py):
    om = _model()
    path = str(tmp_path / "model_bestValidScore")
    om.model.save(path)
    assert checkpoint.file_format(path) == "hdf5"
    f = fake_h5py.File(path, "r")
    assert f.attrs["keras_version"] == b"2.0.4" and f.attrs["backend"] == b"tensorflow"
    mc = json.loads(f.attrs["model_config"].decode())
    assert mc["class_name"] == "Model"
    classes = [l["class_name"] for l in mc["config"]["layers"]]
    # model.py:43-99: data, mask, observed -> concat -> second mask -> concat -> (Dense, Dropout) x 2 -> Dense -> Multiply
    assert classes == ["InputLayer", "InputLayer", "InputLayer", "Concatenate", "InputLayer", "Concatenate",
                       "Dense", "Dropout", "Dense", "Dropout", "Dense", "Multiply"]
    assert [l[0] for l in mc["config"]["input_layers"]] == ["input_1", "input_3", "input_2", "input_4"]   # model.py:89-97
    drop = [l for l in mc["config"]["layers"] if l["class_name"] == "Dropout"][0]["config"]
    assert drop["rate"] == 0.3 and drop["noise_shape"] == [8, 12]                                          # model.py:73
    g = f["model_weights"]
    assert [n.decode() for n in g.attrs["layer_names"]] == ["dense_1", "dense_2", "dense_3"]
    assert [n.decode() for n in g["dense_2"].attrs["weight_names"]] == ["dense_2/kernel:0", "dense_2/bias:0"]
    w = om.model.get_weights()
    assert np.array_equal(np.asarray(g["dense_2"]["dense_2/kernel:0"]), w[2]) and g["dense_3"]["dense_3/bias:0"].shape == (30,)
    assert "optimizer_weights" not in f                         # train.py:183-189 strips it anyway
    back = load_model(path)
    assert back.owner.config() == om.config()
    for a, b in zip(back.get_weights(), w):
        assert a.dtype == np.float32 and np.array_equal(a, b)
    # weights only (model.py:102-107)
    om.save_weights(str(tmp_path / "w"))
    other = _model()
    other.model.set_weights([x * 0 for x in w])
    other.model.load_weights(str(tmp_path / "w"))
    for a, b in zip(other.model.get_weights(), w):
        assert np.array_equal(a, b)


def test_reference_style_donor_file_loads(tmp_path, fake_h5py):
    """A file as Keras itself lays it out: every layer of the graph listed in layer_names (inputs, concatenate,
    dropout, multiply carry no weights), an optimizer_weights group, Keras' own model_config and no ocf_config."""
    rs = np.random.RandomState(0)
    N, H = 20, 7
    kernels = [rs.normal(size=(2 * N, H)).astype(np.float32), rs.normal(size=(H, N)).astype(np.float32)]
    biases = [rs.normal(size=H).astype(np.float32), rs.normal(size=N).astype(np.float32)]
    path = str(tmp_path / "donor_epoch_4_bestValidScore")
    f = fake_h5py.File(path, "w")
    layers = ["input_1", "input_3", "concatenate_1", "dense_1", "dropout_1", "input_2", "dense_2", "multiply_1"]
    mc = {"class_name": "Model", "config": {"layers": [
        {"class_name": "Dense", "name": "dense_1", "config": {"units": H, "activation": "sigmoid",
                                                              "kernel_regularizer": {"class_name": "L1L2", "config": {"l1": 0.0, "l2": 0.001}}}},
        {"class_name": "Dropout", "name": "dropout_1", "config": {"rate": 0.2, "noise_shape": [64, H]}},
        {"class_name": "Dense", "name": "dense_2", "config": {"units": N, "activation": "linear"}}]}}
    f.attrs["model_config"] = json.dumps(mc).encode()
    f.attrs["keras_version"] = b"2.0.4"
    g = f.create_group("model_weights")
    g.attrs["layer_names"] = np.array([n.encode() for n in layers])
    for n in layers:
        lg = g.create_group(n)
        if n.startswith("dense"):
            i = int(n[-1]) - 1
            lg.attrs["weight_names"] = np.array([(n + "/kernel:0").encode(), (n + "/bias:0").encode()])
            for wn, arr in ((n + "/kernel:0", kernels[i]), (n + "/bias:0", biases[i])):
                d = lg.create_dataset(wn, arr.shape, dtype=arr.dtype)
                d[...] = arr
        else:
            lg.attrs["weight_names"] = np.array([], dtype="S1")
    f.create_group("optimizer_weights").create_dataset("Adagrad/accum:0", (3,), dtype=np.float32)
    f.close()
    donor = load_model(path)
    cfg = donor.owner.config()
    assert cfg["numlayers"] == 1 and cfg["num_hidden_units"] == [H] and cfg["input_shape"] == N and cfg["batch_size"] == 64
    assert cfg["dense_activation"] == "sigmoid" and cfg["use_causal_info"] and not cfg["use_both_masks"]
    assert cfg["dropout_probability"] == 0.2 and cfg["l2_weight_regulatization"] == 0.001
    got = donor.get_weights()
    for a, b in zip(got, [kernels[0], biases[0], kernels[1], biases[1]]):
        assert np.array_equal(a, b)
    # the nested-DAE transfer takes it as a donor (model.py:142-170)
    np.random.seed(1)
    deeper = omni_model(3, H, N, 64, dense_activation="sigmoid", use_causal_info=True)
    deeper.load_and_fix_for_denoising_autoencoders(donor)
    w = deeper.model.get_weights()
    assert np.array_equal(w[0], kernels[0]) and np.array_equal(w[6], kernels[1]) and deeper.trainable == [False, True, True, False]


def test_hdf5_file_without_h5py_is_a_loud_error(tmp_path, fake_h5py, monkeypatch):
    om = _model()
    path = str(tmp_path / "m")
    om.model.save(path)
    monkeypatch.setitem(sys.modules, "h5py", None)
    with pytest.raises(RuntimeError, match="h5py"):
        load_model(path)
