"""
Pose normalisation (W1->W2->W3) and dataset loaders against fixtures produced
by the REFERENCE's own pose_utils.py / datasets.py / base_dataset.py (oracle/gen_golden_datasets.py).
Host logic runs on CPU; everything per-ray (process_data, the dataset objects) needs the GPU.
"""
import os

import numpy as np
import pytest

import nerf_tf2_b200 as nb
from nerf_tf2_b200 import pose_utils as pu
from oracle import scene_files

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gp():
    return np.load(os.path.join(GOLD, "ref_pose_utils.npz"))


@pytest.fixture(scope="module")
def gd():
    return np.load(os.path.join(GOLD, "ref_datasets.npz"))


def same(a, b, tol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    if tol == 0.0:
        np.testing.assert_array_equal(a, b)
    else:
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


# ------------------------------------------------------------------ pose_utils, function by function
def test_pose_algebra_bit_identical_to_reference(gp):
    poses, pts, lines = gp["poses"], gp["points"], gp["lines"]
    same(pu.make_4x4(poses[0][:3]), gp["make_4x4"])
    same(pu.make_homogeneous(pts), gp["make_homogeneous"])
    same(pu.normalize(pts[0]), gp["normalize_1d"])
    same(pu.normalize(pts), gp["normalize_2d"])
    same(pu.rotate_vectors(poses[1], pts), gp["rotate_vectors"])
    same(pu.transform_points(poses[2], pts), gp["transform_points_4x4"])
    same(pu.transform_points(poses[2][:3], pts), gp["transform_points_3x4"])
    same(pu.batched_transform_points(poses, pts), gp["batched_transform_points"])
    same(pu.transform_line_segments(poses[3], lines), gp["transform_line_segments"])
    same(pu.batched_transform_line_segments(poses, lines), gp["batched_transform_line_segments"])
    with pytest.raises(ValueError):
        pu.transform_points(np.eye(3), pts)
    with pytest.raises(AssertionError):
        pu.rotate_vectors(np.eye(5), pts)


def test_new_world_frame_bit_identical_to_reference(gp):
    poses = gp["poses"]
    same(pu.solve_min_dist_point(poses), gp["solve_min_dist_point"])
    same(np.stack(pu.compute_new_world_basis(poses)), gp["basis"])
    for om in ("average", "min_dist_solve"):
        same(pu.compute_new_world_origin(poses, om), gp[f"origin_{om}"])
        for bm in ("identity", "compute"):
            same(pu.calculate_new_world_transform(poses, om, bm), gp[f"W1_to_W2_{om}_{bm}"])
    same(pu.reconfigure_poses(poses, gp["W1_to_W2_min_dist_solve_compute"]), gp["reconfigure_poses"])
    with pytest.raises(ValueError):
        pu.compute_new_world_origin(poses, "nope")
    with pytest.raises(ValueError):
        pu.calculate_new_world_transform(poses, "average", "nope")


def test_new_world_frame_properties(gp):
    """W2 is a rigid frame: the transform is orthonormal, the origin is the least-squares point of the
    optical axes, and the SGD prototype walks towards that same point."""
    poses = gp["poses"]
    T = pu.calculate_new_world_transform(poses, "min_dist_solve", "compute")
    R = T[:3, :3]
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-6)
    p = pu.solve_min_dist_point(poses)
    np.testing.assert_allclose(pu.transform_points(T, p[None])[0], 0.0, atol=1e-6)

    def cost(q):
        d = pu.normalize(poses[:, :3, 2])
        diff = q[None] - poses[:, :3, 3]
        return np.sum(np.sum(diff * diff, 1) - np.sum(diff * d, 1) ** 2)
    q = pu.optimize_min_dist_point(poses)
    assert cost(p) <= cost(q) < cost(np.zeros(3))
    # manual basis from a file
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "rot.npy")
        Rm = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
        np.save(path, Rm)
        Tm = pu.calculate_new_world_transform(poses, "average", "manual", manual_rotation=path)
        np.testing.assert_allclose(Tm[:3, :3], Rm.T, atol=1e-12)
        sp = pu.create_spherical_path(3.0, 55.0, 7, path)
        np.testing.assert_allclose(sp[:, :3, 3], (Rm @ gp["spherical_path_r3_i55_n7"][:, :3, 3].T).T, atol=1e-12)


def test_scene_scale_bit_identical_to_reference(gp):
    p2, bounds, Ks = gp["reconfigure_poses"], gp["bounds"], gp["intrinsics"]
    same(pu.get_corner_ray_points(p2, bounds, Ks, 480, 640), gp["corner_ray_points"])
    for bm in ("include_corners", "central_ray"):
        same(np.float64(pu.calculate_scene_scale(p2, bounds, bm, Ks, 480, 640)), gp[f"scene_scale_{bm}"])
    with pytest.raises(ValueError):
        pu.calculate_scene_scale(p2, bounds, "nope", Ks, 480, 640)
    s = float(gp["scene_scale_factor"])
    p3, b3 = pu.reconfigure_scene_scale(p2, bounds, s)
    same(p3, gp["poses_W3"])
    same(b3, gp["bounds_W3"])
    p1, b1 = pu.reconfigure_scene_scale(p2[0], bounds[0], 0.5)
    same(p1, gp["pose_W3_single"])
    same(b1, gp["bounds_W3_single"])
    q, b = pu.reconfigure_scene_scale(p2, bounds, 1.0)          # scale >= 1: untouched, same objects
    assert q is p2 and b is bounds
    # the point of W3: every camera centre and far point lies in [-1,1]^3
    far_pts = p3[:, :3, 3] + b3[:, 1:2] * p3[:, :3, 2]
    assert np.abs(p3[:, :3, 3]).max() <= 1.0 and np.abs(far_pts).max() <= 1.0
    same(pu.create_spherical_path(3.0, 55.0, 7, None), gp["spherical_path_r3_i55_n7"])


def test_scale_imgs_and_intrinsics_matches_reference(gp):
    imgs, Ks = gp["imgs"], gp["intrinsics"][:2]
    si, sk = pu.scale_imgs_and_intrinsics(imgs, Ks, 0.5)
    same(si, gp["imgs_half"])
    same(sk, gp["intrinsics_half"])
    a, b = pu.scale_imgs_and_intrinsics(imgs, Ks, None)
    assert a is imgs and b is Ks


# ------------------------------------------------------------------ loaders (host part)
def _params(kind, root, save_dir, **kw):
    return nb.load_params(scene_files.config_overrides(kind, root, save_dir, **kw))


CASES = [("BlenderDataset", "blender", "wb", True, None), ("BlenderDataset", "blender", "nowb_half", False, 0.5),
         ("CustomDataset", "custom", "std", True, None)]


@pytest.fixture(scope="module")
def scenes(tmp_path_factory):
    root = tmp_path_factory.mktemp("scenes")
    return {"blender": scene_files.write_blender_scene(str(root / "blender")),
            "custom": scene_files.write_custom_scene(str(root / "custom")), "root": str(root)}


@pytest.mark.parametrize("kind,tag,vtag,white_bg,scale_imgs", CASES)
def test_loader_and_reconfigure_match_reference(gd, scenes, kind, tag, vtag, white_bg, scale_imgs):
    save_dir = os.path.join(scenes["root"], f"{tag}_{vtag}_meta")
    params = _params(kind, scenes[tag], save_dir, white_bg=white_bg, scale_imgs=scale_imgs)
    data_splits, num_imgs, obj = nb.get_data_and_metadata_for_splits(params, return_dataset_obj=True)
    assert type(obj).__name__ == kind
    key = f"{tag}_{vtag}"
    for split in ("train", "val", "test"):
        assert num_imgs[split] == int(gd[f"{key}_{split}_num"])
        for field in ("imgs", "poses", "bounds", "intrinsics"):
            same(getattr(data_splits[split], field), gd[f"{key}_{split}_raw_{field}"])
    reconf = obj.validate_and_reconfigure_data(data_splits)
    for split in ("train", "val", "test"):
        for field in ("imgs", "poses", "bounds", "intrinsics"):
            same(getattr(reconf[split], field), gd[f"{key}_{split}_W3_{field}"])
    T, adj = obj.load_reconfig_params()              # written by validate_and_reconfigure_data
    same(T, gd[f"{key}_W1_to_W2"])
    same(adj, gd[f"{key}_adj_scale"])


def test_loader_error_behaviour(scenes):
    root = scenes["blender"]
    cfg = scene_files.config_overrides("BlenderDataset", root, None)
    cfg["blender_dataset"]["val"] = {"num": 2, "frac": 0.5}
    with pytest.raises(ValueError):
        nb.get_dataset_obj(nb.load_params(cfg))
    cfg = scene_files.config_overrides("CustomDataset", scenes["custom"], None)
    cfg["system"]["white_bg"] = True
    with pytest.raises(AssertionError):
        nb.get_dataset_obj(nb.load_params(cfg))
    cfg["system"].update(white_bg=False, dataset_type="Nope")
    with pytest.raises(ValueError):
        nb.get_dataset_obj(nb.load_params(cfg))
    with pytest.raises(AssertionError):
        nb.CustomDataset.camera_model_params_to_intrinsics("FISHEYE", [1, 2, 3])
    bad = np.array([[1.0, 0.1, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    with pytest.raises(AssertionError):
        nb.datasets.Dataset._validate_intrinsic_matrix(bad)
    # mixed image sizes are rejected
    params = _params("BlenderDataset", root, None)
    data_splits, _, obj = nb.get_data_and_metadata_for_splits(params, return_dataset_obj=True)
    d = data_splits["val"]
    data_splits["val"] = d._replace(imgs=d.imgs[:, :-1])
    with pytest.raises(AssertionError):
        obj.validate_and_reconfigure_data(data_splits)


# ------------------------------------------------------------------ per-ray part (device)
@pytest.mark.gpu
@pytest.mark.parametrize("kind,tag,vtag,white_bg,scale_imgs", CASES)
def test_process_data_matches_reference(gd, scenes, kind, tag, vtag, white_bg, scale_imgs):
    save_dir = os.path.join(scenes["root"], f"{tag}_{vtag}_meta_gpu")
    params = _params(kind, scenes[tag], save_dir, white_bg=white_bg, scale_imgs=scale_imgs)
    data_splits, _, obj = nb.get_data_and_metadata_for_splits(params, return_dataset_obj=True)
    reconf = obj.validate_and_reconfigure_data(data_splits)
    rays = obj.process_data(reconf["val"])
    key = f"{tag}_{vtag}"
    assert rays.rays_o.is_cuda
    same(rays.rays_o.cpu().numpy(), gd[f"{key}_val_rays_rays_o"])          # origins: a broadcast copy
    same(rays.near.cpu().numpy(), gd[f"{key}_val_rays_near"])
    same(rays.far.cpu().numpy(), gd[f"{key}_val_rays_far"])
    same(rays.rgb.cpu().numpy(), gd[f"{key}_val_rays_rgb"])
    # directions: fp64 maths rounded to fp32 -- within 1 ulp of the reference's NumPy result
    d, ref = rays.rays_d.cpu().numpy(), gd[f"{key}_val_rays_rays_d"]
    assert np.abs(d - ref).max() <= np.spacing(np.float32(1.0))
    if vtag != "nowb_half":
        same(obj._shuffle(rays).rgb.cpu().numpy(), gd[f"{key}_val_rays_shuffled_rgb"])
    host = obj.process_data(reconf["val"], on_device=False)
    assert isinstance(host.rays_o, np.ndarray) and host.rgb.dtype == np.float32


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["iterate", "sample"])
def test_dataset_objects_feed_the_model(scenes, mode):
    import torch
    save_dir = os.path.join(scenes["root"], f"meta_{mode}")
    params = _params("BlenderDataset", scenes["blender"], save_dir, dataset_mode=mode)
    datasets, num_imgs, img_HW, obj = nb.get_tf_datasets_and_metadata_for_splits(params, return_dataset_obj=True)
    H, W = img_HW
    assert (H, W) == (scene_files.H, scene_files.W) and num_imgs == {"train": 6, "test": 2, "val": 3}
    bs = params.data.batch_size
    # val: ragged last batch kept, rays of 3 images in order
    val = list(datasets["val"])
    assert sum(x[0][0].shape[0] for x in val) == 3 * H * W and val[-1][0][0].shape[0] == (3 * H * W) % bs
    train = list(datasets["train"])
    if mode == "iterate":
        per_epoch = (6 * H * W) // bs
        assert len(train) == per_epoch * params.data.iterate_mode.repeat_count
        assert all(x[0][0].shape[0] == bs for x in train)
        # the two epochs are identical (one fixed shuffle, base_dataset.py:623-651), and skip() advances
        assert torch.equal(train[0][1][0], train[per_epoch][1][0])
        skipped = list(datasets["train"].skip(3))
        assert len(skipped) == len(train) - 3 and torch.equal(skipped[0][0][1], train[3][0][1])
    else:
        assert len(train) == 6 * params.data.sample_mode.repeat_count
        (ro, rd, near, far), (rgb,) = train[0]
        assert ro.shape == (bs, 3) and rgb.shape == (bs, 3) and near.shape == (bs, 1)
        assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0
        # all rays of one batch come from ONE image: a single origin
        assert torch.unique(ro, dim=0).shape[0] == 1
        np.testing.assert_allclose(torch.linalg.norm(rd, dim=1).cpu().numpy(), 1.0, atol=1e-5)
    # the batches go straight into a train step
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, N_coarse=32, N_fine=32), precision="bf16", seed=0)
    x, y = train[0]
    logs = nerf.train_step((x, y))
    assert np.isfinite(list(logs.values())[0])
    # per-view render dataset from a W1 pose through the saved reconfig.npz
    raw, _ = obj.get_data_and_metadata_for_splits()
    c2w, K = raw["test"].poses[0].astype(np.float64), raw["test"].intrinsics[0].astype(np.float64)
    rds = obj.create_dataset_for_render(H, W, c2w, np.array([2.0, 6.0]), K, reconfig_poses=True)
    T, adj = obj.load_reconfig_params()
    want_o = (np.diag([adj, adj, adj, 1.0]) @ T @ c2w)[:3, 3].astype(np.float32)
    first = next(iter(rds))[0]
    np.testing.assert_allclose(first[0][0].cpu().numpy(), want_o, rtol=1e-6)
    np.testing.assert_allclose(float(first[3][0]), 6.0 * float(adj), rtol=1e-6)
    assert sum(b[0][0].shape[0] for b in rds) == H * W


@pytest.mark.gpu
def test_evaluate_split_is_the_eval_script_loop(scenes, tmp_path):
    """main/eval.py over the tiny blender scene: W1 poses through reconfig.npz, per-image PSNR, eval_*.png files."""
    save_dir = os.path.join(scenes["root"], "meta_eval")
    params = _params("BlenderDataset", scenes["blender"], save_dir)
    data_splits, num_imgs, obj = nb.get_data_and_metadata_for_splits(params, return_dataset_obj=True)
    obj.validate_and_reconfigure_data(data_splits)                    # writes reconfig.npz, as a training run would have
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, N_coarse=32, N_fine=32), precision="bf16", seed=1)
    res = nb.render.evaluate_split(nerf, obj, data_splits["test"], save_dir=str(tmp_path))
    assert res["psnr_vals"].shape == (num_imgs["test"],) and np.all(np.isfinite(res["psnr_vals"]))
    assert sorted(os.listdir(tmp_path)) == [f"eval_{i:05d}.png" for i in range(num_imgs["test"])]
    # the first view by hand: same pose chain, same image
    T, adj = obj.load_reconfig_params()
    d = data_splits["test"]
    pose3, b3 = pu.reconfigure_scene_scale(pu.reconfigure_poses(d.poses[0].astype(np.float64), T), d.bounds[0].astype(np.float64), adj)
    H, W = d.imgs[0].shape[:2]
    r = nb.render.render_view(nerf, H, W, pose3, b3, d.intrinsics[0].astype(np.float64), gt_u8=d.imgs[0], depth_maps=False)
    assert abs(r["psnr"] - res["psnr_vals"][0]) < 1e-9
    from PIL import Image
    assert np.array_equal(np.array(Image.open(tmp_path / "eval_00000.png")), r["img_u8"].reshape(H, W, 3).cpu().numpy())


@pytest.mark.gpu
def test_train_script_flow_and_resume(scenes, tmp_path):
    """main/train.py as a function on the tiny blender scene: epochs planned from the repeat count, CustomSaver
    checkpoints after every validation run, and setup_model(params with set_weights) resuming from one of them."""
    import torch
    cfg = scene_files.config_overrides("BlenderDataset", scenes["blender"], os.path.join(scenes["root"], "meta_train"),
                                       dataset_mode="sample")
    cfg["system"].update(steps_per_epoch=3, validation_freq=1, initial_epoch=0)
    cfg["data"]["sample_mode"]["repeat_count"] = 1                      # 6 train images -> 6 steps -> 2 epochs
    cfg["model"] = {"save": {"save_dir": str(tmp_path), "save_optimizer_state": True},
                    "load": {"load_dir": str(tmp_path), "load_tag": "", "set_weights": False, "skip_optimizer": False}}
    cfg["sampling"] = {"N_coarse": 32, "N_fine": 32, "perturb": True, "lin_inv_depth": True}
    params = nb.load_params(cfg)
    assert nb.train.plan_epochs(params, {"train": 6}, (scene_files.H, scene_files.W)) == (6, 3, 2)
    nerf, hist = nb.train.launch(params, precision="bf16", seed=0)
    assert hist.epoch == [0, 1] and len(hist.history["val_psnr_metric"]) == 2 and nerf.optimizer.iterations == 6
    tags = sorted(f[:-len("_coarse.npz")] for f in os.listdir(tmp_path) if f.endswith("_coarse.npz"))
    assert len(tags) == 2 and tags[0].startswith("000000_") and tags[1].startswith("000001_")
    # resume: the reference's setup_model restores weights + optimizer when model.load.set_weights is on
    cfg["model"]["load"].update(load_tag=tags[1], set_weights=True)
    resumed = nb.setup_model(nb.load_params(cfg), precision="bf16", seed=123)
    assert torch.equal(resumed.flat_params, nerf.flat_params) and resumed.optimizer.iterations == 6
    assert torch.equal(resumed.optimizer.m, nerf.optimizer.m)
    cfg["model"]["load"]["skip_optimizer"] = True
    fresh_opt = nb.setup_model(nb.load_params(cfg), precision="bf16", seed=123)
    assert torch.equal(fresh_opt.flat_params, nerf.flat_params) and fresh_opt.optimizer.iterations == 0
