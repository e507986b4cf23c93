"""
Host-side mirror of the reference's `nerf/utils/ray_utils.py` (same function names, argument
meaning and result dictionaries), backed by the sm_100a kernels behind the C ABI
(include/nerfb200.h). Inputs/outputs are CUDA fp32 torch tensors instead of TF tensors.

Differences from the reference, all required by the parity contract:
  * `u_vals` (fixed uniforms) can be passed explicitly to the two samplers; when omitted the
    kernels draw Philox uniforms keyed by (seed, global ray id) -- the reference calls
    `tf.random.uniform` (utils/ray_utils.py:230, :355);
  * `perturb=False` works (the reference raises UnboundLocalError, utils/ray_utils.py:226/263).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, load, ptr, stream_ptr


def _dev(device):
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _f32c(t, device):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    return t.to(device=device, dtype=torch.float32).contiguous()


def _cam_arrays(intrinsic, c2w, ctype, np_dtype):
    K = np.ascontiguousarray(np.asarray(intrinsic, dtype=np_dtype).reshape(9))
    P = np.ascontiguousarray(np.asarray(c2w, dtype=np_dtype).reshape(16))
    return K, P, K.ctypes.data_as(C.POINTER(ctype)), P.ctypes.data_as(C.POINTER(ctype))


def get_rays(H, W, intrinsic, c2w, ray0=0, n_rays=None, device=None):
    """ray_utils.get_rays (utils/ray_utils.py:6-51): fp64 maths, fp32 result [n,3] x2.
    `ray0`/`n_rays` select a contiguous range of the row-major ray ids (ray sharding)."""
    device = _dev(device)
    n = H * W - ray0 if n_rays is None else n_rays
    ro = torch.empty((n, 3), device=device, dtype=torch.float32)
    rd = torch.empty((n, 3), device=device, dtype=torch.float32)
    K, P, kp, pp = _cam_arrays(intrinsic, c2w, C.c_double, np.float64)
    check(load().nerfb200_get_rays(H, W, kp, pp, ray0, n, ptr(ro), ptr(rd), stream_ptr()), "get_rays")
    return ro, rd


def get_rays_tf(H, W, intrinsic, c2w, ray0=0, n_rays=None, device=None):
    """ray_utils.get_rays_tf (utils/ray_utils.py:53-106): fp32 maths."""
    device = _dev(device)
    n = H * W - ray0 if n_rays is None else n_rays
    ro = torch.empty((n, 3), device=device, dtype=torch.float32)
    rd = torch.empty((n, 3), device=device, dtype=torch.float32)
    K, P, kp, pp = _cam_arrays(intrinsic, c2w, C.c_float, np.float32)
    check(load().nerfb200_get_rays_f32(H, W, kp, pp, ray0, n, ptr(ro), ptr(rd), stream_ptr()), "get_rays_tf")
    return ro, rd


def get_rays_at(H, W, intrinsic, c2w, pixel_ids):
    """Rays of selected pixels only (sample-mode training, core/base_dataset.py:555-621)."""
    n = pixel_ids.shape[0]
    ro = torch.empty((n, 3), device=pixel_ids.device, dtype=torch.float32)
    rd = torch.empty((n, 3), device=pixel_ids.device, dtype=torch.float32)
    K, P, kp, pp = _cam_arrays(intrinsic, c2w, C.c_float, np.float32)
    check(load().nerfb200_get_rays_at(H, W, kp, pp, ptr(pixel_ids, torch.int32), n, ptr(ro), ptr(rd),
                                      stream_ptr()), "get_rays_at")
    return ro, rd


def create_depth_map(pred_depth, H, W, scale_factor, map_type, intrinsic=None, C_to_W2=None):
    """ray_utils.create_depth_map (utils/ray_utils.py:108-135). pred_depth: CUDA tensor [H*W]."""
    if map_type == "type_1":
        return (pred_depth * (1 / scale_factor)).reshape(H, W)
    if map_type == "type_2":
        out = torch.empty((H * W,), device=pred_depth.device, dtype=torch.float32)
        K, P, kp, pp = _cam_arrays(intrinsic, C_to_W2, C.c_double, np.float64)
        check(load().nerfb200_depth_type2(H, W, kp, pp, float(scale_factor), ptr(pred_depth.contiguous()),
                                          ptr(out), stream_ptr()), "create_depth_map")
        return out.reshape(H, W)
    raise ValueError(f"Invalid map_type: {map_type}")


def sample_coarse(N_coarse, lin_inv_depth, perturb, near, far, u_vals=None, seed=0, ray0=0, step_state=None):
    """Kernel-level stratified sampler: returns (t_vals[B,Nc], bin_edges[B,Nc+1]). `step_state`: optional device
    int64[2] whose second entry is XORed into the seed on the device (CUDA-graph replays, include/nerfb200.h)."""
    near = near.reshape(-1).contiguous()
    far = far.reshape(-1).contiguous()
    B = near.shape[0]
    t = torch.empty((B, N_coarse), device=near.device, dtype=torch.float32)
    edges = torch.empty((B, N_coarse + 1), device=near.device, dtype=torch.float32)
    check(load().nerfb200_sample_coarse(B, N_coarse, int(bool(lin_inv_depth)), int(bool(perturb)), ptr(near),
                                        ptr(far), ptr(u_vals, allow_none=True), seed,
                                        ptr(step_state, torch.int64, allow_none=True), ray0, ptr(t), ptr(edges),
                                        stream_ptr()), "sample_coarse")
    return t, edges


def make_inputs(rays_o, rays_d, t_vals):
    """xyz_inputs / dir_inputs exactly as the reference materialises them (utils/ray_utils.py:251-258)."""
    B, S = t_vals.shape
    xyz = torch.empty((B * S, 3), device=t_vals.device, dtype=torch.float32)
    dirs = torch.empty((B * S, 3), device=t_vals.device, dtype=torch.float32)
    check(load().nerfb200_make_inputs(B, S, ptr(rays_o), ptr(rays_d), ptr(t_vals), ptr(xyz), ptr(dirs),
                                      stream_ptr()), "make_inputs")
    return xyz, dirs


def positional_encode(x, L):
    """PositionalEncoder.call (core/model.py:305-332) on [R,3] -> [R,3+6L]."""
    R = x.shape[0]
    out = torch.empty((R, 3 + 6 * L), device=x.device, dtype=torch.float32)
    check(load().nerfb200_positional_encode(R, L, ptr(x), ptr(out), stream_ptr()), "positional_encode")
    return out


def create_input_batch_coarse_model(params, rays_o, rays_d, near, far, u_vals=None, seed=0, ray0=0,
                                    materialize=True):
    """ray_utils.create_input_batch_coarse_model (utils/ray_utils.py:137-274)."""
    s = params.sampling
    t_vals, edges = sample_coarse(s.N_coarse, s.lin_inv_depth, s.perturb, near, far, u_vals, seed, ray0)
    left, right = edges[:, :-1], edges[:, 1:]
    data = {
        "bin_data": {"bin_edges": edges, "left_edges": left, "right_edges": right,
                     "bin_widths": right - left},
        "t_vals": t_vals,
    }
    if materialize:
        data["xyz_inputs"], data["dir_inputs"] = make_inputs(rays_o, rays_d, t_vals)
    return data


def sample_fine(N_fine, bin_weights, bin_edges, t_vals_coarse, u_vals=None, seed=0, ray0=0, debug=False, step_state=None):
    """Kernel-level hierarchical sampler. Returns t_sorted[B,Nc+Nf] (and, with debug=True, a dict
    with piece_idxs, the fp32 cdf the indices were searched in, and the unsorted t_fine)."""
    B, Nc = bin_weights.shape
    dev = bin_weights.device
    t_sorted = torch.empty((B, Nc + N_fine), device=dev, dtype=torch.float32)
    idx = cdf = tf = None
    if debug:
        idx = torch.empty((B, N_fine), device=dev, dtype=torch.int32)
        cdf = torch.empty((B, Nc + 1), device=dev, dtype=torch.float32)
        tf = torch.empty((B, N_fine), device=dev, dtype=torch.float32)
    check(load().nerfb200_sample_fine(B, Nc, N_fine, ptr(bin_weights.contiguous()), ptr(bin_edges.contiguous()),
                                      ptr(t_vals_coarse.contiguous()), ptr(u_vals, allow_none=True), seed,
                                      ptr(step_state, torch.int64, allow_none=True), ray0, ptr(t_sorted), ptr(idx, torch.int32, allow_none=True),
                                      ptr(cdf, allow_none=True), ptr(tf, allow_none=True), stream_ptr()),
          "sample_fine")
    if debug:
        return t_sorted, {"piece_idxs": idx, "cdf": cdf, "t_vals_fine": tf}
    return t_sorted


def create_input_batch_fine_model(params, rays_o, rays_d, bin_weights, bin_data, t_vals_coarse, u_vals=None,
                                  seed=0, ray0=0, materialize=True):
    """ray_utils.create_input_batch_fine_model (utils/ray_utils.py:276-406)."""
    t_vals = sample_fine(params.sampling.N_fine, bin_weights, bin_data["bin_edges"], t_vals_coarse, u_vals,
                         seed, ray0)
    data = {"t_vals": t_vals}
    if materialize:
        data["xyz_inputs"], data["dir_inputs"] = make_inputs(rays_o, rays_d, t_vals)
    return data


def post_process_model_output(sample_rgb, sigma, t_vals, white_bg=False, need_weights=True):
    """ray_utils.post_process_model_output (utils/ray_utils.py:484-551), incl. compute_weights /
    sigma_to_alpha (:408-482). sample_rgb [B*S,3], sigma [B*S,1] or [B*S], t_vals [B,S]."""
    B, S = t_vals.shape
    dev = t_vals.device
    weights = torch.empty((B, S), device=dev, dtype=torch.float32) if need_weights else None
    pred_rgb = torch.empty((B, 3), device=dev, dtype=torch.float32)
    pred_depth = torch.empty((B,), device=dev, dtype=torch.float32)
    acc_map = torch.empty((B,), device=dev, dtype=torch.float32)
    check(load().nerfb200_composite_fwd(B, S, ptr(sigma.reshape(-1)), ptr(sample_rgb), ptr(t_vals.contiguous()),
                                        int(bool(white_bg)), ptr(weights, allow_none=True), ptr(pred_rgb),
                                        ptr(pred_depth), ptr(acc_map), stream_ptr()), "composite_fwd")
    out = {"acc_map": acc_map, "pred_rgb": pred_rgb, "pred_depth": pred_depth}
    if need_weights:
        out["weights"] = weights
    return out


def compute_weights(sigma, t_vals, N_samples=None):
    """ray_utils.compute_weights (utils/ray_utils.py:426-482)."""
    B, S = t_vals.shape
    rgb = torch.zeros((B * S, 3), device=t_vals.device, dtype=torch.float32)
    return post_process_model_output(rgb, sigma, t_vals, False)["weights"]


def sigma_to_alpha(sigma, diffs):
    """ray_utils.sigma_to_alpha (utils/ray_utils.py:408-424); elementwise, provided for API parity
    (the integrator kernel fuses it)."""
    return 1 - torch.exp(-sigma * diffs)


def composite_train(sample_rgb, sigma, t_vals, white_bg, rgb_gt, B_global, loss, metric_state=None, need_weights=True):
    """The integrator of a training step in ONE launch (nerfb200_composite_train): post_process_model_output, the
    MeanSquaredError term of this output added to `loss` (and the PSNRMetric state to `metric_state`), and the gradient
    w.r.t. (sigma, sample_rgb). Returns (post_proc dict, d_sigma [B*S], d_rgb [B*S,3])."""
    B, S = t_vals.shape
    dev = t_vals.device
    weights = torch.empty((B, S), device=dev, dtype=torch.float32) if need_weights else None
    pred_rgb = torch.empty((B, 3), device=dev, dtype=torch.float32)
    pred_depth = torch.empty((B,), device=dev, dtype=torch.float32)
    acc_map = torch.empty((B,), device=dev, dtype=torch.float32)
    d_sigma = torch.empty((B * S,), device=dev, dtype=torch.float32)
    d_rgb = torch.empty((B * S, 3), device=dev, dtype=torch.float32)
    check(load().nerfb200_composite_train(B, S, ptr(sigma.reshape(-1)), ptr(sample_rgb), ptr(t_vals.contiguous()),
                                          int(bool(white_bg)), ptr(rgb_gt.contiguous()), int(B_global),
                                          ptr(weights, allow_none=True), ptr(pred_rgb), ptr(pred_depth), ptr(acc_map),
                                          ptr(d_sigma), ptr(d_rgb), ptr(loss), ptr(metric_state, allow_none=True),
                                          stream_ptr()), "composite_train")
    out = {"acc_map": acc_map, "pred_rgb": pred_rgb, "pred_depth": pred_depth}
    if need_weights:
        out["weights"] = weights
    return out, d_sigma, d_rgb


def composite_backward(sample_rgb, sigma, t_vals, white_bg, d_pred_rgb):
    """Gradient of post_process_model_output's pred_rgb w.r.t. (sigma, sample_rgb)."""
    B, S = t_vals.shape
    d_sigma = torch.empty((B * S,), device=t_vals.device, dtype=torch.float32)
    d_rgb = torch.empty((B * S, 3), device=t_vals.device, dtype=torch.float32)
    check(load().nerfb200_composite_bwd(B, S, ptr(sigma.reshape(-1)), ptr(sample_rgb), ptr(t_vals.contiguous()),
                                        int(bool(white_bg)), ptr(d_pred_rgb.contiguous()), ptr(d_sigma), ptr(d_rgb),
                                        stream_ptr()), "composite_bwd")
    return d_sigma, d_rgb
