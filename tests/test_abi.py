"""CPU: the C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

import nerf_tf2_b200 as nb
from nerf_tf2_b200 import _lib


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "nerfb200.h")).read()
    return sorted(set(re.findall(r"NERFB200_API[^;]*?\b(nerfb200_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 22
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nerfb200.h but not exported"
    # the ctypes binding covers exactly the declared surface
    assert sorted(_lib.SIGNATURES) == syms


def test_abi_version_and_param_layout():
    lib = _lib.load()
    assert lib.nerfb200_abi_version() == 3
    offs = _lib.param_offsets()
    assert offs[0] == 0 and offs[1] == 63 * 256 and offs[-1] == _lib.PARAMS_PER_MODEL == 595844
    from oracle import model as om
    sizes = []
    for ln in om.LAYER_NAMES:
        fi, fo = om.LAYER_SHAPES[ln]
        sizes += [fi * fo, fo]
    assert offs == list(np.concatenate([[0], np.cumsum(sizes)]))


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    # all of these fail argument checks before any CUDA call
    assert lib.nerfb200_composite_fwd(4, 1, None, None, None, 0, None, None, None, None, None) == 10001
    assert b"S" in lib.nerfb200_last_error()
    assert lib.nerfb200_sample_fine(4, 48, 128, 1, 1, 1, None, 0, None, 0, 1, None, None, None, None) == 10002
    assert lib.nerfb200_sample_fine(4, 64, 100, 1, 1, 1, None, 0, None, 0, 1, None, None, None, None) == 10002
    assert lib.nerfb200_mlp_forward(None, 0, 1, 1, None, None, None, None, None, None, 1, None, None, None) == 10001
    assert lib.nerfb200_get_rays(0, 4, None, None, 0, 0, None, None, None) == 10001
    # the phase-split backward validates like mlp_backward: no context -> EINVAL, before any CUDA call
    assert lib.nerfb200_mlp_backward_data(None, 0, 1, 1, None, None, None, 1, None, None, 0, None) == 10001
    assert lib.nerfb200_mlp_backward_weights(None, 0, 1, 1, None, 1, None, None, 0, None) == 10001
    assert b"context" in lib.nerfb200_last_error()
    # the one-call ray march validates its context, shapes and precision before touching the device
    fwd = lambda ctx, B, prec: lib.nerfb200_forward(ctx, B, 64, 128, 1, 1, 1, None, None, None, None, None, None, 0, None, 0, None, prec,
                                                    None, None, None, None, None, None, None, None, None, None)
    assert fwd(None, 4, 1) == 10001 and b"context" in lib.nerfb200_last_error()
    assert lib.nerfb200_forward_workspace_bytes(1000, 64, 128) >= 4 * 1000 * (64 + 65 + 4 * 64 + 64 + 192 + 4 * 192)
    assert lib.nerfb200_set_option(None, 1, 1) == 10001
    assert lib.nerfb200_step_advance(None, None) == 10001
    assert lib.nerfb200_mlp_workspace_bytes(1000, 1, 0) == 0
    # the peer-memory gradient exchange validates world / rank / size before it allocates anything
    h = ctypes.c_void_p()
    assert lib.nerfb200_peer_create(9, 0, 1024, ctypes.byref(h)) == 10001 and b"world" in lib.nerfb200_last_error()
    assert lib.nerfb200_peer_create(2, 2, 1024, ctypes.byref(h)) == 10001
    assert lib.nerfb200_peer_create(2, 0, 1023, ctypes.byref(h)) == 10001 and b"multiple of 4" in lib.nerfb200_last_error()
    assert not h.value
    assert lib.nerfb200_peer_allreduce(None, None) == 10001
    assert lib.nerfb200_peer_allreduce_adam(None, 4, None, None, None, 0, None, None) == 10001
    assert lib.nerfb200_peer_destroy(None) == 0
    assert lib.nerfb200_mlp_stash_bytes(10, 0) == 10 * (63 + 27 + 8 * 256 + 256 + 128 + 3 + 1) * 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nb.NerfB200Error):
        nb.NeRF(nb.make_params())
    with pytest.raises(nb.NerfB200Error):
        nb.ray_utils.get_rays(4, 4, np.eye(3), np.eye(4))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nerf-tf2_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f
