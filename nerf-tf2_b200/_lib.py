"""
ctypes binding of libnerfb200.so (C ABI in include/nerfb200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing, or a
compute entry point is called without a CUDA device, this module raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnerfb200.so")

FP32, BF16, FP16, TF32 = 0, 1, 2, 3
PRECISIONS = {"fp32": FP32, "bf16": BF16, "fp16": FP16, "tf32": TF32}
PACK_BIT = {BF16: 1, FP16: 2, TF32: 4}          # NERFB200_OPT_PACK_MASK bits
OPT_PRECISE_LAST, OPT_PACK_MASK, OPT_DEBUG = 1, 2, 3
COARSE, FINE = 0, 1
PARAMS_PER_MODEL = 595844
PARAMS_TOTAL = 2 * PARAMS_PER_MODEL

_i64, _i32, _vp, _u64, _dbl = C.c_int64, C.c_int, C.c_void_p, C.c_uint64, C.c_double

# name -> (restype, argtypes); must list EVERY symbol declared in include/nerfb200.h
SIGNATURES = {
    "nerfb200_last_error": (C.c_char_p, []),
    "nerfb200_abi_version": (_i32, []),
    "nerfb200_launch_count": (_i64, []),
    "nerfb200_param_offsets": (_i32, [C.POINTER(_i64)]),
    "nerfb200_get_rays": (_i32, [_i32, _i32, C.POINTER(_dbl), C.POINTER(_dbl), _i64, _i64, _vp, _vp, _vp]),
    "nerfb200_get_rays_f32": (_i32, [_i32, _i32, C.POINTER(C.c_float), C.POINTER(C.c_float), _i64, _i64, _vp, _vp, _vp]),
    "nerfb200_get_rays_at": (_i32, [_i32, _i32, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _i64, _vp, _vp, _vp]),
    "nerfb200_sample_coarse": (_i32, [_i64, _i32, _i32, _i32, _vp, _vp, _vp, _u64, _vp, _i64, _vp, _vp, _vp]),
    "nerfb200_make_inputs": (_i32, [_i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_positional_encode": (_i32, [_i64, _i32, _vp, _vp, _vp]),
    "nerfb200_create": (_i32, [C.POINTER(_vp)]),
    "nerfb200_destroy": (_i32, [_vp]),
    "nerfb200_pack_weights": (_i32, [_vp, _vp, _vp]),
    "nerfb200_set_option": (_i32, [_vp, _i32, _i32]),
    "nerfb200_mlp_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "nerfb200_mlp_stash_bytes": (_i64, [_i64, _i32]),
    "nerfb200_mlp_forward": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "nerfb200_mlp_backward": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "nerfb200_mlp_backward_data": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp]),
    "nerfb200_mlp_backward_weights": (_i32, [_vp, _i32, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _vp]),
    "nerfb200_forward_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "nerfb200_forward": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _i64, _vp, _i32, _vp,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_composite_fwd": (_i32, [_i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_composite_bwd": (_i32, [_i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "nerfb200_composite_train": (_i32, [_i64, _i32, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_sample_fine": (_i32, [_i64, _i32, _i32, _vp, _vp, _vp, _vp, _u64, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_mse_loss_grad": (_i32, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_adam_step": (_i32, [_i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nerfb200_step_advance": (_i32, [_vp, _vp]),
    "nerfb200_peer_create": (_i32, [_i32, _i32, _i64, C.POINTER(_vp)]),
    "nerfb200_peer_attach": (_i32, [_i32, _i32, _i64, C.POINTER(_vp), _vp, C.POINTER(_vp)]),
    "nerfb200_peer_buffer": (_i32, [_vp, C.POINTER(_vp)]),
    "nerfb200_peer_handle": (_i32, [_vp, C.c_char_p]),
    "nerfb200_peer_connect": (_i32, [_vp, C.c_char_p]),
    "nerfb200_peer_set_timeout": (_i32, [_vp, _i32]),
    "nerfb200_peer_allreduce": (_i32, [_vp, _vp]),
    "nerfb200_peer_allreduce_adam": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nerfb200_peer_status": (_i32, [_vp, C.POINTER(_i32)]),
    "nerfb200_peer_profile": (_i32, [_vp, C.POINTER(C.c_ulonglong)]),
    "nerfb200_peer_disconnect": (_i32, [_vp]),
    "nerfb200_peer_destroy": (_i32, [_vp]),
    "nerfb200_depth_type2": (_i32, [_i32, _i32, C.POINTER(_dbl), C.POINTER(_dbl), _dbl, _vp, _vp, _vp]),
    "nerfb200_sample_pixels": (_i32, [_i64, _i64, _u64, _u64, _vp, _vp]),
    "nerfb200_gather_rgb_u8": (_i32, [_i64, _vp, _vp, _vp, _vp]),
    "nerfb200_postprocess_rgb": (_i32, [_i64, _vp, _vp, _vp, _vp, _vp]),
}


class NerfB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Loads the shared library (once) and binds every symbol. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfB200Error(
            f"{LIB_PATH} is missing: build it with `python nerf-tf2_b200/build.py` "
            "(there is no CPU or PyTorch fallback for the ray-march path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().nerfb200_last_error().decode("utf-8", "replace")
        raise NerfB200Error(f"{what} failed (code {rc}): {msg}")


def require_cuda():
    if not torch.cuda.is_available():
        raise NerfB200Error("a CUDA device (B200, sm_100a) is required: there is no CPU fallback")


def stream_ptr(device=None):
    """The current torch stream of `device` (default: the current device) as a cudaStream_t."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t, dtype=torch.float32, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (validated)."""
    if t is None:
        if allow_none:
            return C.c_void_p(0)
        raise NerfB200Error("NULL tensor")
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise NerfB200Error(f"expected a contiguous CUDA {dtype} tensor, got {t.dtype} "
                            f"{'cuda' if t.is_cuda else 'cpu'} contiguous={t.is_contiguous()}")
    return C.c_void_p(t.data_ptr())


def param_offsets():
    arr = (_i64 * 25)()
    check(load().nerfb200_param_offsets(arr), "param_offsets")
    return list(arr)
