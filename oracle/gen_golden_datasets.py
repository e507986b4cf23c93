"""
TEST INFRASTRUCTURE ONLY -- generates tests/golden/ref_pose_utils.npz and ref_datasets.npz.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden_datasets

Every array in the two fixtures is an output of the reference's OWN code, imported unmodified from
/root/reference: `nerf/utils/pose_utils.py` on seeded random poses, and the loaders of
`nerf/core/datasets.py` + `Dataset.validate_and_reconfigure_data / process_data`
(`nerf/core/base_dataset.py`) run over the tiny on-disk scenes written by oracle/scene_files.py.
Stand-ins needed to import them here: the NumPy TensorFlow shim (tf is imported but not used on these
paths), an `imageio.imread` that opens the file with PIL (the reference asks imageio for its "PNG-PIL"
plugin, i.e. PIL), and an attribute-dict in place of python-box.
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np

from . import scene_files, tf_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


def to_attr(d):
    return AttrDict({k: to_attr(v) for k, v in d.items()}) if isinstance(d, dict) else d


def random_poses(rng, n):
    """Rigid camera->world poses with orthonormal rotations, [n,4,4] float64."""
    out = []
    for _ in range(n):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        m = np.eye(4)
        m[:3, :3] = q
        m[:3, 3] = rng.normal(size=3) * 3.0
        out.append(m)
    return np.array(out)


def gen_pose_utils(pu):
    rng = np.random.default_rng(424242)
    poses = random_poses(rng, 9)
    # inward-facing variant: z axis towards a common point, so min_dist_solve is well conditioned
    for p in poses:
        z = -p[:3, 3] + 0.2 * rng.normal(size=3)
        z /= np.linalg.norm(z)
        x = np.cross([0.1, 0.2, 1.0], z)
        x /= np.linalg.norm(x)
        p[:3, 0], p[:3, 1], p[:3, 2] = x, np.cross(z, x), z
    bounds = np.stack([rng.uniform(0.5, 2.0, 9), rng.uniform(4.0, 7.0, 9)], axis=1)
    Ks = np.array([[[500.0 + 7 * i, 0, 320.5], [0, 480.0 + 3 * i, 239.5], [0, 0, 1]] for i in range(9)])
    pts = rng.normal(size=(17, 3))
    lines = rng.normal(size=(5, 2, 3))
    out = {"poses": poses, "bounds": bounds, "intrinsics": Ks, "points": pts, "lines": lines}
    out["make_4x4"] = pu.make_4x4(poses[0][:3])
    out["make_homogeneous"] = pu.make_homogeneous(pts)
    out["normalize_1d"] = pu.normalize(pts[0])
    out["normalize_2d"] = pu.normalize(pts)
    out["rotate_vectors"] = pu.rotate_vectors(poses[1], pts)
    out["transform_points_4x4"] = pu.transform_points(poses[2], pts)
    out["transform_points_3x4"] = pu.transform_points(poses[2][:3], pts)
    out["batched_transform_points"] = pu.batched_transform_points(poses, pts)
    out["transform_line_segments"] = pu.transform_line_segments(poses[3], lines)
    out["batched_transform_line_segments"] = pu.batched_transform_line_segments(poses, lines)
    out["solve_min_dist_point"] = pu.solve_min_dist_point(poses)
    for om in ("average", "min_dist_solve"):
        out[f"origin_{om}"] = pu.compute_new_world_origin(poses, om)
        for bm in ("identity", "compute"):
            out[f"W1_to_W2_{om}_{bm}"] = pu.calculate_new_world_transform(poses, om, bm)
    xb, yb, zb = pu.compute_new_world_basis(poses)
    out["basis"] = np.stack([xb, yb, zb])
    T = out["W1_to_W2_min_dist_solve_compute"]
    p2 = pu.reconfigure_poses(poses, T)
    out["reconfigure_poses"] = p2
    out["corner_ray_points"] = pu.get_corner_ray_points(p2, bounds, Ks, 480, 640)
    for bm in ("include_corners", "central_ray"):
        out[f"scene_scale_{bm}"] = np.float64(pu.calculate_scene_scale(p2, bounds, bm, Ks, 480, 640))
    s = 0.85 * out["scene_scale_include_corners"]
    p3, b3 = pu.reconfigure_scene_scale(p2, bounds, s)
    out["scene_scale_factor"], out["poses_W3"], out["bounds_W3"] = np.float64(s), p3, b3
    p1, b1 = pu.reconfigure_scene_scale(p2[0], bounds[0], 0.5)      # non-batched
    out["pose_W3_single"], out["bounds_W3_single"] = p1, b1
    imgs = rng.integers(0, 256, size=(2, 24, 36, 3), dtype=np.uint8)
    si, sk = pu.scale_imgs_and_intrinsics(imgs, Ks[:2], 0.5)
    out["imgs"], out["imgs_half"], out["intrinsics_half"] = imgs, si, sk
    out["spherical_path_r3_i55_n7"] = pu.create_spherical_path(3.0, 55.0, 7, None)
    np.savez_compressed(os.path.join(OUT, "ref_pose_utils.npz"), **out)


def gen_datasets(ref_datasets):
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for kind, tag, writer in (("BlenderDataset", "blender", scene_files.write_blender_scene),
                                  ("CustomDataset", "custom", scene_files.write_custom_scene)):
            root = writer(os.path.join(tmp, tag))
            variants = [("wb" if kind == "BlenderDataset" else "std", True, None)]
            if kind == "BlenderDataset":
                variants.append(("nowb_half", False, 0.5))
            for vtag, white_bg, scale_imgs in variants:
                save_dir = os.path.join(tmp, f"{tag}_{vtag}_meta")
                params = to_attr(scene_files.config_overrides(kind, root, save_dir, white_bg=white_bg,
                                                              scale_imgs=scale_imgs))
                data_splits, num_imgs, obj = ref_datasets.get_data_and_metadata_for_splits(params, return_dataset_obj=True)
                reconf = obj.validate_and_reconfigure_data(data_splits)
                T, adj = obj.load_reconfig_params()
                key = f"{tag}_{vtag}"
                out[f"{key}_W1_to_W2"], out[f"{key}_adj_scale"] = T, adj
                for split in ("train", "val", "test"):
                    out[f"{key}_{split}_num"] = num_imgs[split]
                    for field in ("imgs", "poses", "bounds", "intrinsics"):
                        out[f"{key}_{split}_raw_{field}"] = getattr(data_splits[split], field)
                        out[f"{key}_{split}_W3_{field}"] = getattr(reconf[split], field)
                rays = obj.process_data(reconf["val"])
                for field in rays._fields:
                    out[f"{key}_val_rays_{field}"] = getattr(rays, field)
                if vtag in ("wb", "std"):
                    shuffled = obj._shuffle(rays)
                    out[f"{key}_val_rays_shuffled_rgb"] = shuffled.rgb
    np.savez_compressed(os.path.join(OUT, "ref_datasets.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    imageio = types.ModuleType("imageio")

    def imread(path, fmt=None):
        from PIL import Image
        return np.array(Image.open(path))
    imageio.imread = imread
    saved = sys.modules.get("imageio")
    sys.modules["imageio"] = imageio
    try:
        with tf_shim.reference_ray_utils() as (ru, pu, tf):
            gen_pose_utils(pu)
            ref_datasets = importlib.import_module("nerf.core.datasets")
            gen_datasets(ref_datasets)
    finally:
        if saved is None:
            sys.modules.pop("imageio", None)
        else:
            sys.modules["imageio"] = saved
    for f in ("ref_pose_utils.npz", "ref_datasets.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
