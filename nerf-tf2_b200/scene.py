"""
Synthetic 360-degree inward-facing scenes of the reference's shape (no dataset files are
available offline): camera poses on a sphere looking at the origin (the construction of
utils/pose_utils.py:770-840 create_spherical_path), SIMPLE_PINHOLE intrinsics
[1111.111, 400, 400] at 800x800 scaled proportionally (params/config.yaml:63-72), blender
bounds 2.0/6.0 (core/datasets.py:195) and the W2->W3 scene scale applied the way
create_dataset_for_render does (core/base_dataset.py:826-861).
"""
import numpy as np


def _normalize(v):
    return v / (np.sqrt(np.sum(v ** 2, axis=-1, keepdims=True)) + 1e-8)


def spherical_poses(radius=4.0, inclination=40.0, num_cameras=8):
    """Camera-to-world 4x4 poses (Classic-CV: +z looks at the origin), float64 [N,4,4]."""
    az = np.radians(np.linspace(0, 360, num_cameras, endpoint=False, dtype=np.float64))
    inc = np.radians(np.full_like(az, inclination))
    r = np.full_like(az, radius)
    origin = np.stack([r * np.sin(inc) * np.cos(az), r * np.sin(inc) * np.sin(az), r * np.cos(inc)], axis=1)
    z = _normalize(-origin)
    x = _normalize(np.stack([-r * np.sin(inc) * np.sin(az), r * np.sin(inc) * np.cos(az), np.zeros_like(az)], axis=1))
    y = _normalize(np.cross(z, x))
    poses = np.zeros((num_cameras, 4, 4), dtype=np.float64)
    poses[:, :3, 0], poses[:, :3, 1], poses[:, :3, 2], poses[:, :3, 3] = x, y, z, origin
    poses[:, 3, 3] = 1.0
    return poses


def intrinsic_for(H, W):
    f = 1111.111 * (W / 800.0)
    return np.array([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def scale_pose_and_bounds(pose, bounds, scale):
    """pose_utils.reconfigure_scene_scale (utils/pose_utils.py:427-463)."""
    if scale >= 1:
        return pose, bounds
    T = np.eye(4) * scale
    T[3, 3] = 1
    return T @ pose, np.asarray(bounds, dtype=np.float64) * scale


class SyntheticScene:
    """`num_cameras` views of an HxW camera orbiting the origin; scene scaled into [-1,1]^3."""

    def __init__(self, H, W, num_cameras=8, radius=4.0, inclination=40.0, bounds=(2.0, 6.0),
                 adj_scale_factor=0.2125):
        self.H, self.W = H, W
        self.K = intrinsic_for(H, W)
        self.adj_scale_factor = adj_scale_factor
        raw = spherical_poses(radius, inclination, num_cameras)
        self.poses, self.bounds = [], None
        for p in raw:
            q, b = scale_pose_and_bounds(p, np.asarray(bounds, dtype=np.float64), adj_scale_factor)
            self.poses.append(q)
            self.bounds = b
        self.near, self.far = float(np.float32(self.bounds[0])), float(np.float32(self.bounds[1]))

    def __len__(self):
        return len(self.poses)


# Analytic ground truth for end-to-end runs (there are no dataset files offline): a handful of shaded,
# coloured spheres on a white background, ray traced in NumPy from any camera->W3 pose.
SPHERES = np.array([   # cx, cy, cz, radius, r, g, b   (W3 units: the scene lives in [-0.5, 0.5]^3)
    [0.00, 0.00, 0.00, 0.130, 0.85, 0.20, 0.15],
    [0.18, 0.06, 0.03, 0.075, 0.15, 0.55, 0.85],
    [-0.15, 0.12, -0.06, 0.085, 0.20, 0.75, 0.25],
    [0.03, -0.18, 0.07, 0.060, 0.90, 0.80, 0.10],
], dtype=np.float64)


def render_spheres(H, W, K, c2w, spheres=SPHERES, light=(0.3, 0.5, 0.8)):
    """uint8 [H,W,3] image of the sphere scene seen through pinhole K from pose c2w (pixel (u,v) ->
    direction ((u-cx)/fx, (v-cy)/fy, 1) like get_rays, utils/ray_utils.py:6-51)."""
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    d = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], axis=-1).reshape(-1, 3)
    d = _normalize(d @ c2w[:3, :3].T)
    o = c2w[:3, 3]
    L = _normalize(np.asarray(light, dtype=np.float64))
    best_t = np.full(d.shape[0], np.inf)
    rgb = np.ones((d.shape[0], 3))
    for cx, cy, cz, r, cr, cg, cb in spheres:
        oc = o - np.array([cx, cy, cz])
        b = d @ oc
        disc = b * b - (oc @ oc - r * r)
        t = -b - np.sqrt(np.maximum(disc, 0.0))
        hit = (disc > 0) & (t > 0) & (t < best_t)
        n = _normalize((o + t[:, None] * d) - np.array([cx, cy, cz]))
        shade = 0.35 + 0.65 * np.maximum(n @ L, 0.0)
        rgb[hit] = (np.array([cr, cg, cb])[None, :] * shade[:, None])[hit]
        best_t = np.where(hit, t, best_t)
    return np.clip(rgb * 255.0 + 0.5, 0, 255).astype(np.uint8).reshape(H, W, 3)
