"""
TEST INFRASTRUCTURE ONLY -- synthetic 360-degree inward-facing views for the oracle.

Follows the reference's render path: poses from create_spherical_path
(utils/pose_utils.py:770-840; defaults params/config.yaml:50-59), SIMPLE_PINHOLE
intrinsics [1111.111, 400, 400] at 800x800 (params/config.yaml:63-72) scaled
proportionally, blender bounds 2.0/6.0 (core/datasets.py:195), then the
W2->W3 scene scale of create_dataset_for_render (core/base_dataset.py:826-861,
utils/pose_utils.py:reconfigure_scene_scale) and get_rays in fp64 cast to fp32.
"""
import numpy as np

from . import ray_march as rm

F32 = np.float32


def intrinsic_for(H, W):
    f = 1111.111 * (W / 800.0)
    return np.array([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def reconfigure_scene_scale(pose, bounds, s):
    if s >= 1:
        return pose, bounds
    T = np.eye(4) * s
    T[3, 3] = 1
    return T @ pose.copy(), bounds.copy() * s


def synthetic_view(H, W, view=0, num_cameras=8, radius=4.0, inclination=40.0,
                   bounds=(2.0, 6.0), adj_scale_factor=0.2125):
    poses = rm.create_spherical_path(radius, inclination, num_cameras)
    K = intrinsic_for(H, W)
    pose, b = reconfigure_scene_scale(poses[view], np.asarray(bounds, dtype=np.float64),
                                      adj_scale_factor)
    rays_o, rays_d = rm.get_rays(H, W, K, pose)
    n = rays_d.shape[0]
    near = np.full((n, 1), b[0], dtype=np.float64).astype(F32)
    far = np.full((n, 1), b[1], dtype=np.float64).astype(F32)
    return {"rays_o": np.ascontiguousarray(rays_o.astype(F32)),
            "rays_d": np.ascontiguousarray(rays_d.astype(F32)),
            "near": near, "far": far, "K": K, "c2w": pose, "bounds": b,
            "adj_scale_factor": adj_scale_factor}
