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
                ray0=0, n_rays=None, c2w_W2=None, lazy_psnr=False):
    """Renders one view (or the ray range [ray0, ray0+n_rays) of it). `c2w`/`bounds` are the W3 (scene-scaled) pose
    and bounds the rays are marched with. Returns a dict of DEVICE tensors: img_u8 [n,3], acc_map [n],
    depth_type_1 [n], depth_type_2 [n] (full views only) and `psnr` (float, eval.py definition) when `gt_u8`
    ([H,W,3] or [n,3] uint8) is given. The depth maps are in W2 units like main/render.py:103-112: type_1 is
    depth / scale_factor, type_2 the camera-space z of `o + d * depth / scale_factor` for the UNSCALED camera->W2
    pose -- `c2w_W2`, or `c2w` with the scene scale undone when it is not given."""
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
        out["sq_err"] = sq           # fp64 sum of squared errors of the clipped 8-bit image (device)
        if not lazy_psnr:            # lazy: the caller turns sq_err into a PSNR later, without a sync per view
            out["psnr"] = psnr_from_sq_err(float(sq.item()), n)
    if depth_maps and scale_factor is not None:
        out["depth_type_1"] = fine["pred_depth"] * (1 / scale_factor)
        if ray0 == 0 and n == H * W:
            if c2w_W2 is None:      # reconfigure_scene_scale applied diag(s,s,s,1) only when s < 1 (pose_utils.py:427-463)
                c2w_W2 = np.asarray(c2w, dtype=np.float64)
                if scale_factor < 1:
                    c2w_W2 = np.diag([1 / scale_factor] * 3 + [1.0]) @ c2w_W2
            out["depth_type_2"] = ray_utils.create_depth_map(fine["pred_depth"], H, W, scale_factor, "type_2",
                                                             intrinsic, c2w_W2).reshape(-1)
    return out


def psnr_from_sq_err(sq_err, n_rays):
    """main/eval.py:57-61 + core/ops.py:249-257: -10 log10(mean over n_rays*3 of the squared error)."""
    mse = sq_err / (n_rays * 3)
    return float("inf") if mse == 0 else -10.0 * math.log10(mse)


def render_view_sharded(nerf, H, W, c2w, bounds, intrinsic, process_group=None, dst=0, rays=None):
    """One view marched by ALL ranks of `process_group`: rank r takes the contiguous ray range
    dist.shard_range(H*W, r, G), replicated weights, no communication on the data path; the only collective is the final
    gather of the image rows. Returns, on rank `dst` (every rank when dst is None), the device tensor [H*W, 5] fp32 =
    (r, g, b, depth, acc) of the fine model; None elsewhere. `rays` = this rank's shard (rays_o, rays_d, near, far)
    when the rays are already resident; otherwise the get_rays kernel generates exactly the shard. The Philox streams
    are keyed by the global ray id, so the image does not depend on G."""
    import torch.distributed as tdist
    world = tdist.get_world_size(process_group) if (tdist.is_available() and tdist.is_initialized()) else 1
    rank = tdist.get_rank(process_group) if world > 1 else 0
    a, b = nbdist.shard_range(H * W, rank, world)
    dev = nerf.device
    if rays is None:
        ro, rd = ray_utils.get_rays(H, W, intrinsic, c2w, a, b - a, device=dev)
        near = torch.full((b - a, 1), float(np.float32(bounds[0])), device=dev)
        far = torch.full((b - a, 1), float(np.float32(bounds[1])), device=dev)
    else:
        ro, rd, near, far = rays
        assert ro.shape[0] == b - a, "rays must be this rank's shard_range of the view"
    _, fine = nerf.render_rays(ro, rd, near, far, ray0=a, need_weights=False, keep_coarse=False)
    rows = torch.cat([fine["pred_rgb"], fine["pred_depth"].reshape(-1, 1), fine["acc_map"].reshape(-1, 1)], dim=1)
    if world == 1:
        return rows
    return nbdist.gather_rows(rows, H * W, process_group, dst)


def predict_view_sharded(nerf, ds_host_shard, n_total, process_group=None, dst=0):
    """End-to-end form of render_view_sharded for HOST rays: `ds_host_shard` is a RayDataset over this rank's
    shard_range of the view's rays in host memory (NumPy). NeRF.predict stages them to the device through pinned
    buffers and marches them; the fine model's (rgb, depth, acc) rows are gathered to rank `dst` and copied to the host
    there. Returns the [n_total, 5] NumPy array on `dst`, None elsewhere."""
    import torch.distributed as tdist
    world = tdist.get_world_size(process_group) if (tdist.is_available() and tdist.is_initialized()) else 1
    _, fine = nerf.predict(ds_host_shard, return_weights=False, as_numpy=False)
    rows = torch.cat([fine["pred_rgb"], fine["pred_depth"].reshape(-1, 1), fine["acc_map"].reshape(-1, 1)], dim=1)
    full = rows if world == 1 else nbdist.gather_rows(rows, n_total, process_group, dst)
    if full is None:
        return None
    pin = nerf._ws.get(("pin_rows", n_total))
    if pin is None:
        pin = nerf._ws[("pin_rows", n_total)] = torch.empty((n_total, 5), dtype=torch.float32).pin_memory()
    pin.copy_(full, non_blocking=True)
    torch.cuda.current_stream(nerf.device).synchronize()
    return pin.numpy()


def evaluate_views(nerf, H, W, poses, bounds, intrinsic, gts_u8, scale_factor=None, process_group=None, depth_maps=False):
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
        r = render_view(nerf, H, W, poses[i], b, intrinsic, gt_u8=gts_u8[i], scale_factor=scale_factor, depth_maps=depth_maps,
                        lazy_psnr=True)
        local[i:i + 1].copy_(r["sq_err"])          # stays on the device: no host sync between views
    if world > 1:
        tdist.all_reduce(local, group=process_group)      # the job's only collective: 8 bytes per view
    vals = np.array([psnr_from_sq_err(float(v), H * W) for v in local.cpu().numpy()], dtype=np.float64)
    return {"psnr_vals": vals, "mean_psnr": float(vals.mean()), "last_view_psnr": float(vals[-1])}


def render_spherical_path(nerf, render_params, adj_scale_factor, save_dir=None):
    """The loop of main/render.py:51-117: `num_cameras` poses on a sphere (create_spherical_path, in W2), intrinsics from
    the COLMAP camera model, bounds = 0.25/0.75 of the diameter unless given, every view scaled to W3 with
    `adj_scale_factor` (what create_dataset_for_render does with reconfig_poses=False) and ray-marched on the device.
    Returns a list of dicts with HOST arrays `img_u8 [H,W,3]`, `depth_type_1 [H,W]`, `depth_type_2 [H,W]`,
    `acc_map [H,W]`; with `save_dir`, also writes rgb/*.png and depth_type_1|depth_type_2|acc_map/*.npy under the
    script's names (render_00000.png ...)."""
    from . import pose_utils
    from .datasets import CustomDataset
    poses = pose_utils.create_spherical_path(radius=render_params.radius, num_cameras=render_params.num_cameras,
                                             inclination=render_params.inclination,
                                             manual_rotation=render_params.manual_rotation)
    K = CustomDataset.camera_model_params_to_intrinsics(render_params.camera_model_name, render_params.camera_model_params)
    if render_params.bounds is None:
        diameter = 2 * render_params.radius
        bounds = np.array([0.25 * diameter, 0.75 * diameter], dtype=np.float64)
    else:
        bounds = np.array(render_params.bounds, dtype=np.float64)
    H, W = render_params.img_size
    s = float(adj_scale_factor)
    zfill = int(np.log10(render_params.num_cameras) + 5)
    dirs = {}
    if save_dir is not None:
        import os
        for sub in ("rgb", "depth_type_1", "depth_type_2", "acc_map"):
            dirs[sub] = os.path.join(save_dir, sub)
            os.makedirs(dirs[sub], exist_ok=True)
    frames = []
    for i in range(render_params.num_cameras):
        pose3, bounds3 = pose_utils.reconfigure_scene_scale(poses[i], bounds, s)
        r = render_view(nerf, H, W, pose3, bounds3, K, scale_factor=s, c2w_W2=poses[i])
        out = {"img_u8": r["img_u8"].reshape(H, W, 3).cpu().numpy(),
               "depth_type_1": r["depth_type_1"].reshape(H, W).cpu().numpy(),
               "depth_type_2": r["depth_type_2"].reshape(H, W).cpu().numpy(),
               "acc_map": r["acc_map"].reshape(H, W).cpu().numpy()}
        frames.append(out)
        if save_dir is not None:
            import os
            from PIL import Image
            name = f"render_{str(i).zfill(zfill)}"
            Image.fromarray(out["img_u8"]).save(os.path.join(dirs["rgb"], f"{name}.png"))
            for k in ("depth_type_1", "depth_type_2", "acc_map"):
                np.save(os.path.join(dirs[k], f"{name}.npy"), out[k])
    return frames


def evaluate_split(nerf, dataset_obj, data, save_dir=None):
    """The loop of main/eval.py:26-67 over one split's SceneLevelData (`data`, poses still in W1 as the loaders return
    them): each pose goes W1 -> W2 -> W3 through the saved reconfig.npz (what create_dataset_for_render does with
    reconfig_poses=True), the view is ray-marched and scored against its image with the script's PSNR. Returns
    {"psnr_vals", "mean_psnr", "last_view_psnr"} (the script logs np.mean of the LAST psnr only) and,
    with `save_dir`, writes eval_00000.png ... like the script."""
    from . import pose_utils
    T, adj = dataset_obj.load_reconfig_params()
    H, W = data.imgs[0].shape[:2]
    zfill = int(np.log10(len(data.imgs)) + 5)
    if save_dir is not None:
        import os
        os.makedirs(save_dir, exist_ok=True)
    vals = []
    for i in range(len(data.imgs)):
        K = np.asarray(data.intrinsics[i], dtype=np.float64)
        dataset_obj._validate_intrinsic_matrix(K=K)
        pose2 = pose_utils.reconfigure_poses(np.asarray(data.poses[i], dtype=np.float64), T)
        pose3, bounds3 = pose_utils.reconfigure_scene_scale(pose2, np.asarray(data.bounds[i], dtype=np.float64), adj)
        r = render_view(nerf, H, W, pose3, bounds3, K, gt_u8=data.imgs[i], depth_maps=False)
        vals.append(r["psnr"])
        if save_dir is not None:
            import os
            from PIL import Image
            Image.fromarray(r["img_u8"].reshape(H, W, 3).cpu().numpy()).save(os.path.join(save_dir, f"eval_{str(i).zfill(zfill)}.png"))
    vals = np.asarray(vals, dtype=np.float64)
    return {"psnr_vals": vals, "mean_psnr": float(vals.mean()), "last_view_psnr": float(vals[-1])}
