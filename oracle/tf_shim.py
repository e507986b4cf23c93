"""
TEST INFRASTRUCTURE ONLY -- not part of the product path.

A NumPy fp32 stand-in for the handful of `tensorflow` ops that the reference's
`nerf/utils/ray_utils.py` calls, so that file can be imported and executed
*unmodified* from /root/reference in the build container (TensorFlow 2.6 is not
installable here). Used by `oracle/gen_golden.py` to
produce the committed fixtures under tests/golden/ and by the in-container tests
that pin `oracle/ray_march.py` against the real reference code. Never imported by
the package, by -m gpu tests, by smoke() or by bench.py (the reference tree does
not exist on the GPU box).

Semantics follow SURVEY.md Appendix C (TF-2.6 op semantics).
"""
import contextlib
import importlib
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"

F32 = np.float32


def _f32(x):
    if isinstance(x, np.ndarray):
        return x
    return np.asarray(x, dtype=F32)


class _Random:
    """tf.random.uniform replacement: pops pre-seeded arrays from a queue."""

    def __init__(self):
        self.queue = []

    def uniform(self, shape, dtype=None, **kw):
        shape = tuple(int(s) for s in shape)
        assert self.queue, "tf_shim: no preset uniforms queued for tf.random.uniform"
        u = self.queue.pop(0)
        assert u.shape == shape, (u.shape, shape)
        assert u.dtype == F32
        return u


def _linspace(start, stop, num, axis=0):
    # TF 2.6 math_ops.linspace_nd: endpoints exact, interior = start + delta * i.
    start = _f32(start)
    stop = _f32(stop)
    num = int(num)
    es = np.expand_dims(start, axis)
    ee = np.expand_dims(stop, axis)
    delta = (ee - es) / F32(num - 1)
    idx = np.arange(1, num - 1, dtype=F32)
    shp = [1] * es.ndim
    shp[axis] = -1
    res = es + delta * idx.reshape(shp)
    return np.concatenate([es, res, ee], axis=axis).astype(F32)


def _cumprod(x, axis=0, exclusive=False):
    x = np.asarray(x)
    out = np.cumprod(x, axis=axis, dtype=x.dtype)  # sequential in dtype
    if exclusive:
        out = np.roll(out, 1, axis=axis)
        sl = [slice(None)] * x.ndim
        sl[axis] = slice(0, 1)
        out[tuple(sl)] = 1
    return out


def _searchsorted(seq, values, side="left"):
    out = np.empty(values.shape, dtype=np.int32)
    for b in range(seq.shape[0]):
        out[b] = np.searchsorted(seq[b], values[b], side=side)
    return out


def _gather(params, indices, axis=None, batch_dims=0):
    assert axis == 1 and batch_dims == 1
    return np.take_along_axis(params, indices.astype(np.int64), axis=1)


def make_tf_module():
    tf = types.ModuleType("tensorflow")
    tf.float32 = F32
    tf.random = _Random()
    tf.range = lambda start=0, limit=None, dtype=None: np.arange(start, limit, dtype=dtype)
    tf.meshgrid = lambda *a, indexing="xy": np.meshgrid(*a, indexing=indexing)
    tf.ones_like = np.ones_like
    tf.zeros_like = np.zeros_like
    tf.stack = lambda vals, axis=0: np.stack(vals, axis=axis)
    tf.reshape = lambda x, shape: np.reshape(x, tuple(int(s) for s in shape))
    tf.transpose = np.transpose
    tf.sqrt = np.sqrt
    tf.exp = np.exp
    tf.maximum = lambda a, b: np.maximum(a, np.asarray(b, dtype=a.dtype))
    tf.reduce_sum = lambda x, axis=None, keepdims=False: np.sum(x, axis=axis, keepdims=keepdims, dtype=x.dtype)
    tf.broadcast_to = lambda x, shape: np.broadcast_to(x, tuple(int(s) for s in shape))
    tf.shape = lambda x: tuple(np.shape(x))
    tf.squeeze = lambda x, axis=None: np.squeeze(x, axis=axis)
    tf.concat = lambda vals, axis=0: np.concatenate(vals, axis=axis)
    tf.fill = lambda dims, value: np.full(tuple(int(d) for d in dims), value, dtype=F32)
    tf.where = lambda c, a, b: np.where(c, a, b)
    tf.stop_gradient = lambda x: x
    tf.sort = lambda x, axis=-1: np.sort(x, axis=axis)
    tf.cumsum = lambda x, axis=0: np.cumsum(x, axis=axis, dtype=x.dtype)  # sequential fp32
    tf.gather = _gather
    tf.searchsorted = _searchsorted
    tf.linspace = _linspace
    tf.math = types.SimpleNamespace(cumprod=_cumprod)
    tf.linalg = types.SimpleNamespace(matmul=lambda a, b: np.matmul(a, b))
    return tf


@contextlib.contextmanager
def reference_ray_utils():
    """
    Yields (ray_utils, pose_utils, tf) with `ray_utils`/`pose_utils` being the
    reference's own modules imported from /root/reference over the shim.
    """
    tf = make_tf_module()
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "nerf", "nerf.utils",
                                             "nerf.utils.ray_utils", "nerf.utils.pose_utils")}
    sys.modules["tensorflow"] = tf
    for k in list(sys.modules):
        if k == "nerf" or k.startswith("nerf."):
            del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ray_utils = importlib.import_module("nerf.utils.ray_utils")
        pose_utils = importlib.import_module("nerf.utils.pose_utils")
        yield ray_utils, pose_utils, tf
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in list(sys.modules):
            if k == "nerf" or k.startswith("nerf."):
                del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)


class Params:
    """Attribute-access stand-in for the python-box `params` object."""

    def __init__(self, N_coarse=64, N_fine=128, perturb=True, lin_inv_depth=True,
                 white_bg=True, batch_size=4096):
        self.sampling = types.SimpleNamespace(N_coarse=N_coarse, N_fine=N_fine,
                                              perturb=perturb, lin_inv_depth=lin_inv_depth)
        self.system = types.SimpleNamespace(white_bg=white_bg, run_eagerly=False, log_images=False)
        self.data = types.SimpleNamespace(batch_size=batch_size)
