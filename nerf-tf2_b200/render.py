"""
Per-view render driver: what `main/render.py:79-117` and `main/eval.py:42-64`
do around `nerf.predict`, kept on the device so that only finished images cross PCIe.

  reference (per view)                                   here
  ------------------------------------------------------ ---------------------------------------------
  create_dataset_for_render: get_rays (NumPy, fp64)       get_rays kernel (fp64 maths) into HBM
  nerf.predict -> both dicts incl. [N,S] weights to host  render_rays, fine dict only, no weights
  np.clip(pred_rgb*255, 0, 255).astype(uint8)             postprocess_rgb kernel -> uint8 [H,W,3]
  psnr_metric_numpy(gt/255, clip(pred*255)/255)           same kernel, fp64 accumulator
  create_depth_map type_1 / type_2                        elementwise / depth_type2 kernel
"""
import ctypes as C
import math

import numpy as np
import torch

from . import dist as nbdist, ray_utils
from ._lib import check, load, ptr, stream_ptr


def render_view(nerf, H, W, c2w, bounds, intrinsic, gt_u8=None, scale_factor=None, depth_maps=True,
                ray0=0, n_rays=None):
    """Renders one view (or the ray range [ray0, ray0+n_rays) of it). Returns a dict of DEVICE tensors:
    img_u8 [n,3], acc_map [n], depth_type_1 [n], depth_type_2 [n] (full views only) and `psnr` (float,
    eval.py definition) when `gt_u8` ([H,W,3] or [n,3] uint8) is given."""
    dev = nerf.device
    n = H * W - ray0 if n_rays is None else n_rays
    ro, rd = ray_utils.get_rays(H, W, intrinsic, c2w, ray0, n, device=dev)
    near = torch.full((n, 1), float(np.float32(bounds[0])), device=dev)
    far = torch.full((n, 1), float(np.float32(bounds[1])), device=dev)
    _, fine = nerf.render_rays(ro, rd, near, far, ray0=ray0, need_weights=False, keep_coarse=False)
    out = {"pred_rgb": fine["pred_rgb"], "acc_map": fine["acc_map"], "pred_depth": fine["pred_depth"]}
    img = torch.empty((n, 3), device=dev, dtype=torch.uint8)
    gt = None
    sq = None
    if gt_u8 is not None:
        gt = torch.as_tensor(gt_u8, dtype=torch.uint8).reshape(-1, 3).to(dev).contiguous()
        assert gt.shape[0] == n, "ground truth must cover exactly the rendered rays"
        sq = torch.zeros(1, device=dev, dtype=torch.float64)
    check(load().nerfb200_postprocess_rgb(n * 3, ptr(fine["pred_rgb"]), ptr(gt, torch.uint8, allow_none=True),
                                          ptr(img, torch.uint8), ptr(sq, torch.float64, allow_none=True), stream_ptr()),
          "postprocess_rgb")
    out["img_u8"] = img
    if sq is not None:
        out["sq_err"] = sq
        mse = float(sq.item()) / (n * 3)
        out["psnr"] = float("inf") if mse == 0 else -10.0 * math.log10(mse)
    if depth_maps and scale_factor is not None:
        out["depth_type_1"] = fine["pred_depth"] * (1 / scale_factor)
        if ray0 == 0 and n == H * W:
            out["depth_type_2"] = ray_utils.create_depth_map(fine["pred_depth"], H, W, scale_factor, "type_2",
                                                             intrinsic, c2w).reshape(-1)
    return out


def evaluate_views(nerf, H, W, poses, bounds, intrinsic, gts_u8, scale_factor=None, process_group=None):
    """eval.py-shaped test-set loop (BASELINE config 4): renders every view, returns per-view PSNRs and
    their mean. Views are dealt round-robin to the ranks of `process_group` (no collective on the data
    path; the PSNR values are gathered at the end). Note: the reference logs np.mean(psnr) -- the LAST
    image's PSNR (main/eval.py:66); the mean over all views is returned here and the
    last-view value is available as result["last_view_psnr"]."""
    import torch.distributed as tdist
    world = tdist.get_world_size(process_group) if (tdist.is_available() and tdist.is_initialized()) else 1
    rank = tdist.get_rank(process_group) if world > 1 else 0
    mine = nbdist.shard_views(len(poses), rank, world)
    local = torch.zeros(len(poses), device=nerf.device, dtype=torch.float64)
    for i in mine:
        b = bounds[i] if np.ndim(bounds) == 2 else bounds
        r = render_view(nerf, H, W, poses[i], b, intrinsic, gt_u8=gts_u8[i], scale_factor=scale_factor, depth_maps=False)
        local[i] = r["psnr"]
    if world > 1:
        tdist.all_reduce(local, group=process_group)
    vals = local.cpu().numpy()
    return {"psnr_vals": vals, "mean_psnr": float(vals.mean()), "last_view_psnr": float(vals[-1])}
