"""
h5lite -- the minimal HDF5 reader/writer behind Keras `.h5` weight files.
The reader is checked against a file written by the real libhdf5 (the MATLAB v7.3 sample that ships
with SciPy's test data: user block, superblock v0, symbol-table group, v1 object headers, attribute,
contiguous float64 dataset); the writer by round trip and structure. The GPU test drives the
CustomSaver / set_everything cycle with `.h5` files.
"""
import os
import struct

import numpy as np
import pytest

import nerf_tf2_b200 as nb
from nerf_tf2_b200 import checkpoint, h5lite


def _libhdf5_sample():
    import scipy.io.matlab
    p = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("SciPy test data (a real libhdf5 file) is not installed")
    return p


def test_reader_on_a_real_libhdf5_file():
    root = h5lite.read_file(_libhdf5_sample())
    assert root.keys() == ["testdouble"]
    ds = root["testdouble"]
    assert ds.value.dtype == np.float64 and ds.value.shape == (9, 1)
    np.testing.assert_allclose(ds.value[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
    assert ds.attrs["MATLAB_class"] == b"double"
    assert "testdouble" in root and "nope" not in root


def _layers(rng, model="coarse"):
    out = []
    shapes = {"dense_0": (63, 256), "dense_5": (319, 256), "dense_9": (283, 128), "rgb": (128, 3), "sigma": (256, 1)}
    for ln in checkpoint.keras_layers(model):
        short = ln.split("/")[1]
        if short.startswith("dense") or short in ("rgb", "sigma"):
            fi, fo = shapes.get(short, (256, 256))
            out.append((ln, [(f"{ln}/kernel:0", rng.normal(size=(fi, fo)).astype(np.float32)),
                             (f"{ln}/bias:0", rng.normal(size=(fo,)).astype(np.float32))]))
        else:
            out.append((ln, []))
    return out


def test_keras_weight_file_round_trip(tmp_path):
    layers = _layers(np.random.default_rng(5))
    path = str(tmp_path / "000001_20.00_coarse.h5")
    h5lite.save_keras_weights(path, layers)
    root = h5lite.read_file(path)
    # the structure Keras writes: root attrs, one (nested) group per layer, datasets under the weight name
    assert [n.decode() for n in root.attrs["layer_names"]] == [ln for ln, _ in layers]
    assert root.attrs["backend"] == b"tensorflow" and root.attrs["keras_version"] == b"2.7.0"
    assert root["coarse/enc_xyz"].attrs["weight_names"].shape == (0,)
    assert [w.decode() for w in root["coarse/dense_5"].attrs["weight_names"]] == ["coarse/dense_5/kernel:0", "coarse/dense_5/bias:0"]
    assert root["coarse/dense_5/coarse/dense_5/kernel:0"].value.shape == (319, 256)
    got = h5lite.load_keras_weights(path)
    want = [w for _, ws in layers for w in ws]
    assert [n for n, _ in got] == [n for n, _ in want] and len(got) == 24
    for (_, a), (_, b) in zip(got, want):
        assert a.dtype == np.float32 and np.array_equal(a, b)
    # Keras' order puts the two heads last, rgb before sigma
    assert [n for n, _ in got][-4:] == ["coarse/rgb/kernel:0", "coarse/rgb/bias:0", "coarse/sigma/kernel:0", "coarse/sigma/bias:0"]


def test_written_file_structure_follows_the_format_spec(tmp_path):
    path = str(tmp_path / "t.h5")
    h5lite.write_file(path, {"g": ("group", {"x": ("dataset", np.arange(6, dtype=np.int32).reshape(2, 3), {"unit": np.array(b"m")}),
                                             "y": ("dataset", np.linspace(0, 1, 5), {})}, {"note": np.array([b"ab", b"cd"])}),
                             "z": ("dataset", np.array([1.5, -2.0], dtype=np.float16), {})}, {"top": np.array(7, dtype=np.int64)})
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8      # superblock v0, 8-byte offsets
    eof = struct.unpack_from("<Q", raw, 40)[0]
    assert eof == len(raw)
    for sig in (b"TREE", b"SNOD", b"HEAP"):
        assert raw.count(sig) == 2                                                              # root group + /g
    root = h5lite.read_file(path)
    assert root.attrs["top"] == 7 and root.keys() == ["g", "z"]
    assert np.array_equal(root["g/x"].value, np.arange(6, dtype=np.int32).reshape(2, 3)) and root["g/x"].attrs["unit"] == b"m"
    assert np.array_equal(root["g/y"].value, np.linspace(0, 1, 5)) and root["g/y"].value.dtype == np.float64
    assert np.array_equal(root["z"].value, np.array([1.5, -2.0], dtype=np.float16))
    assert list(root["g"].attrs["note"]) == [b"ab", b"cd"]
    assert [p for p, _ in root.visit_datasets()] == ["g/x", "g/y", "z"]


def test_unsupported_files_fail_loudly(tmp_path):
    p = tmp_path / "bad.h5"
    p.write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(h5lite.H5Error, match="signature"):
        h5lite.read_file(str(p))
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + b"\0" * 200)                            # superblock v2 = libver latest
    with pytest.raises(h5lite.H5Error, match="superblock version 2"):
        h5lite.read_file(str(p))
    with pytest.raises(h5lite.H5Error):
        h5lite.write_file(str(p), {f"c{i}": ("dataset", np.zeros(1), {}) for i in range(40)})    # > one symbol node


class _FakeVar:
    def __init__(self, name, arr):
        self.name, self._a, self.shape = name, arr, arr.shape

    def numpy(self):
        return self._a


class _FakeSub:
    def __init__(self, name, arrays):
        self.name = name
        self.trainable_variables = [_FakeVar(n, a) for n, a in arrays]
        self.loaded = None

    def set_weights(self, arrays):
        self.loaded = arrays


def test_sub_model_h5_save_load_host_side(tmp_path):
    rng = np.random.default_rng(2)
    arrays = [(n[:-2], a) for _, ws in _layers(rng, "fine") for n, a in ws]
    sub = _FakeSub("fine", arrays)
    path = str(tmp_path / "w_fine.h5")
    checkpoint.save_weights_h5(path, sub)
    checkpoint.load_weights(path, sub)
    assert all(np.array_equal(a, b) for a, (_, b) in zip(sub.loaded, arrays))
    # a file of the other shape family is rejected, as Keras does
    bad = _FakeSub("fine", arrays[:-2] + [("fine/sigma/kernel", np.zeros((128, 1), np.float32)), arrays[-1]])
    with pytest.raises(AssertionError, match="shape"):
        checkpoint.load_weights(path, bad)


@pytest.mark.gpu
def test_checkpoint_round_trip_h5(tmp_path):
    import torch
    a = nb.setup_model(nb.make_params({"system": {"white_bg": True}}), precision="bf16", seed=4)
    saver = nb.CustomSaver(str(tmp_path), weights_format="h5")
    saver.set_model(a)
    saver.on_epoch_end(3, {"psnr_metric": 10.0, "val_psnr_metric": 9.5})
    tag = "000003_9.50"
    assert sorted(os.listdir(tmp_path)) == [f"{tag}_coarse.h5", f"{tag}_fine.h5", f"{tag}_logs.npz", f"{tag}_optimizer.npz"]
    b = nb.setup_model(nb.make_params({"system": {"white_bg": True}}), precision="bf16", seed=77)
    assert not torch.equal(a.flat_params, b.flat_params)
    b.set_everything(str(tmp_path), tag)
    assert torch.equal(a.flat_params, b.flat_params)
    # SubModel.save_weights / load_weights, the Keras entry points
    a.fine_model.save_weights(str(tmp_path / "fine_only.h5"))
    c = nb.setup_model(nb.make_params(), precision="bf16", seed=5)
    c.fine_model.load_weights(str(tmp_path / "fine_only.h5"))
    n = a.flat_params.numel() // 2
    assert torch.equal(c.flat_params[n:], a.flat_params[n:]) and not torch.equal(c.flat_params[:n], a.flat_params[:n])
