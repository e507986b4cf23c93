"""
Pose algebra and the W1 -> W2 -> W3 scene normalisation, the host-side step
that runs ONCE per dataset before any ray reaches the hot path. Same function names, argument meaning
and results as the reference's utils/pose_utils.py (file:line cited per function); float64 NumPy, as
there. O(#cameras) work: it stays on the host on purpose.

Coordinate systems (reference docs): W1 = whatever the dataset's poses use; W2 = W1 re-centred on the
point closest to all optical axes and re-oriented so +y is the mean camera y axis; W3 = W2 scaled so the
points the networks see fall into [-1, 1]^3 (what the positional encoding assumes).
"""
import numpy as np

_EPS = 1e-8


# ---------------------------------------------------------------- 4x4 algebra (pose_utils.py:5-229)
def make_4x4(arr):
    """[3,4] -> [4,4] with last row (0,0,0,1) (pose_utils.py:5-24)."""
    arr = np.asarray(arr)
    assert arr.shape == (3, 4)
    out = np.eye(4)
    out[:3] = arr
    return out


def make_homogeneous(points):
    """[N,3] -> [N,4], last column ones (pose_utils.py:26-42)."""
    assert points.shape[1] == 3
    return np.concatenate([points, np.ones((points.shape[0], 1), dtype=np.float64)], axis=1)


def normalize(vec):
    """v / (|v| + 1e-8) for one vector [D] or a batch [N,D] (pose_utils.py:44-75)."""
    assert vec.ndim <= 2
    mag = np.sqrt(np.sum(vec ** 2, axis=-1))
    return vec / ((mag if vec.ndim == 1 else mag[:, None]) + _EPS)


def rotate_vectors(arr, vectors):
    """(R @ v^T)^T with R = arr[:3,:3]; arr is [3,3], [3,4] or [4,4] (pose_utils.py:77-106)."""
    assert arr.shape in ((3, 3), (3, 4), (4, 4)), "Shape of arr is invalid. Must be either (3, 3), (3, 4) or (4, 4)"
    assert vectors.shape[1] == 3
    return (arr[:3, :3] @ vectors.T).T


def transform_points(arr, points):
    """Rigid/affine transform of [N,3] points by a [3,4] or [4,4] matrix (pose_utils.py:108-138)."""
    assert points.shape[1] == 3
    if arr.shape == (3, 4):
        arr = make_4x4(arr)
    elif arr.shape != (4, 4):
        raise ValueError("Shape of arr is invalid. Must be either (3, 4) or (4, 4)")
    return (arr @ make_homogeneous(points).T).T[:, :3]


def batched_transform_points(arrs, points):
    """M transforms [M,4,4] applied to the same [N,3] points -> [M,N,3] (pose_utils.py:140-171)."""
    assert arrs.shape[1:] == (4, 4), "Shape of arrs is invalid. Must be (M, 4, 4)"
    assert points.shape[1] == 3
    return (arrs @ make_homogeneous(points).T).transpose(0, 2, 1)[..., :3]


def transform_line_segments(arr, lines):
    """[N,2,3] segments, both end points transformed (pose_utils.py:173-200)."""
    assert lines.shape[1:] == (2, 3)
    return transform_points(arr, lines.reshape(-1, 3)).reshape(lines.shape[0], 2, 3)


def batched_transform_line_segments(arrs, lines):
    """[M,4,4] x [N,2,3] -> [M,N,2,3] (pose_utils.py:202-229)."""
    assert lines.shape[1:] == (2, 3)
    return batched_transform_points(arrs, lines.reshape(-1, 3)).reshape(arrs.shape[0], lines.shape[0], 2, 3)


# ---------------------------------------------------------------- W1 -> W2 (pose_utils.py:231-290, 585-768)
def solve_min_dist_point(poses):
    """Least-squares point closest to all optical axes: sum_i (d_i d_i^T - I) p = sum_i (d_i d_i^T - I) o_i
    (pose_utils.py:636-672). Accumulated camera by camera in the reference's order."""
    origins = poses[:, :3, 3]
    dirs = normalize(poses[:, :3, 2])
    A, b = np.zeros((3, 3)), np.zeros((3, 1))
    for o, d in zip(origins, dirs):
        P = d[:, None] @ d[None, :] - np.eye(3)
        A += P
        b += P @ o[:, None]
    return np.squeeze(np.linalg.solve(A, b))


def optimize_min_dist_point(poses, steps=1000, learning_rate=1e-4):
    """The reference's prototype (pose_utils.py:585-634): 1000 plain-SGD steps (lr 1e-4) on
    sum_i |p - o_i|^2 - ((p - o_i).d_i)^2 from p = 0. The gradient is written out instead of taped."""
    origins = poses[:, :3, 3]
    dirs = normalize(poses[:, :3, 2])
    p = np.zeros((1, 3), dtype=np.float64)
    for _ in range(steps):
        diff = p - origins
        along = np.sum(diff * dirs, axis=1, keepdims=True)
        grad = np.sum(2.0 * diff - 2.0 * along * dirs, axis=0, keepdims=True)
        p = p - learning_rate * grad
    return np.squeeze(p.T)


def compute_new_world_origin(poses, method):
    """W2 origin expressed in W1 (pose_utils.py:674-716)."""
    if method == "average":
        return np.mean(poses[:, :3, 3], axis=0)
    if method == "min_dist_solve":
        return solve_min_dist_point(poses)
    if method == "min_dist_opt":
        return optimize_min_dist_point(poses)
    raise ValueError(f"Invalid method: {method}")


def compute_new_world_basis(poses):
    """W2 axes in W1: y = mean camera y; z = x_cam0 x y; x = y x z, each normalised (pose_utils.py:718-768)."""
    y = normalize(np.mean(poses[:, :3, 1], axis=0))
    z = normalize(np.cross(normalize(poses[0, :3, 0]), y))
    x = normalize(np.cross(y, z))
    return x, y, z


def calculate_new_world_transform(poses, origin_method, basis_method, manual_rotation=None):
    """The 4x4 W1 -> W2 transform = inverse of [x y z origin] (pose_utils.py:231-290)."""
    origin = compute_new_world_origin(poses, method=origin_method)
    if basis_method == "identity":
        x, y, z = np.eye(3)
    elif basis_method == "compute":
        x, y, z = compute_new_world_basis(poses)
    elif basis_method == "manual":
        assert manual_rotation is not None
        R = np.load(manual_rotation)
        assert R.shape == (3, 3)
        x, y, z = R[:, 0], R[:, 1], R[:, 2]
    else:
        raise ValueError(f"Invalid basis_method: {basis_method}")
    return np.linalg.inv(make_4x4(np.stack([x, y, z, origin], axis=1)))


def reconfigure_poses(old_poses, W1_to_W2_transform):
    """camera->W1 poses ([N,4,4] or [4,4]) become camera->W2 (pose_utils.py:400-425)."""
    return W1_to_W2_transform @ old_poses


# ---------------------------------------------------------------- W2 -> W3 (pose_utils.py:292-398, 427-463, 520-583)
def get_corner_ray_points(poses, bounds, intrinsics, height, width):
    """Far-plane points of the 4 image-corner rays of every camera, [4N,3] in W2 (pose_utils.py:520-583)."""
    u = np.array([[0, 0, width, width]], dtype=np.float64)
    v = np.array([[0, height, 0, height]], dtype=np.float64)
    x = (u - intrinsics[:, 0, 2, None]) / intrinsics[:, 0, 0, None]
    y = (v - intrinsics[:, 1, 2, None]) / intrinsics[:, 1, 1, None]
    dirs = np.stack([x, y, np.ones_like(x)], axis=-1)
    far = bounds[:, 1:2]
    out = []
    for i, c2w in enumerate(poses):
        d = normalize(rotate_vectors(c2w, normalize(dirs[i])))
        out.append(c2w[None, :3, 3] + far[i] * d)
    return np.array(out).reshape(-1, 3)


def calculate_scene_scale(poses, bounds, bounds_method, intrinsics=None, height=None, width=None):
    """1 / (largest |coordinate| over camera centres, far points on the optical axes and, for
    "include_corners", the far points of the corner rays) (pose_utils.py:292-398)."""
    rays_o = poses[:, :3, 3]
    pts = [rays_o, rays_o + bounds[:, 1][:, None] * poses[:, :3, 2]]
    if bounds_method == "include_corners":
        assert intrinsics is not None and height is not None and width is not None
        pts.append(get_corner_ray_points(poses, bounds, intrinsics, height, width))
    elif bounds_method != "central_ray":
        raise ValueError(f"Invalid bounds_method: {bounds_method}")
    return 1 / np.abs(np.concatenate(pts, axis=0)).max(axis=0).max()


def reconfigure_scene_scale(old_poses, old_bounds, scene_scale_factor):
    """diag(s,s,s,1) @ pose and s * bounds when s < 1, unchanged otherwise (pose_utils.py:427-463)."""
    if scene_scale_factor >= 1:
        return old_poses, old_bounds
    S = np.eye(4) * scene_scale_factor
    S[3, 3] = 1
    return S @ old_poses.copy(), old_bounds.copy() * scene_scale_factor


def scale_imgs_and_intrinsics(old_imgs, old_intrinsics, scale_factor):
    """Optional training-resolution change: INTER_AREA resize of every image and fx,cx,fy,cy scaled
    with it (pose_utils.py:465-518). `None` = untouched."""
    if scale_factor is None:
        return old_imgs, old_intrinsics
    import cv2
    imgs, Ks = [], []
    for img, K in zip(old_imgs, old_intrinsics):
        # the reference converts RGB->BGR->RGB around the resize; INTER_AREA is channel-wise, so that is a no-op
        imgs.append(cv2.resize(np.ascontiguousarray(img), dsize=None, fx=scale_factor, fy=scale_factor,
                               interpolation=cv2.INTER_AREA))
        K = K.copy()
        K[0, 0], K[0, 2] = K[0, 0] * scale_factor, K[0, 2] * scale_factor
        K[1, 1], K[1, 2] = K[1, 1] * scale_factor, K[1, 2] * scale_factor
        Ks.append(K)
    return np.array(imgs), np.array(Ks)


# ---------------------------------------------------------------- render path (pose_utils.py:770-840)
def create_spherical_path(radius, inclination, num_cameras, manual_rotation=None):
    """`num_cameras` camera->world poses on a circle of constant inclination on a sphere around the
    world origin, +z looking at the origin; optional extra rotation from a .npy (pose_utils.py:770-840)."""
    az = np.radians(np.linspace(0, 360, num_cameras, endpoint=False, dtype=np.float64))
    r = np.full_like(az, radius)
    inc = np.radians(np.full_like(az, inclination))
    origin = np.stack([r * np.sin(inc) * np.cos(az), r * np.sin(inc) * np.sin(az), r * np.cos(inc)], axis=1)
    z = normalize(-1 * origin)
    tx, ty = -1 * r * np.sin(inc) * np.sin(az), r * np.sin(inc) * np.cos(az)
    x = normalize(np.stack([tx, ty, np.zeros_like(tx)], axis=1))
    y = normalize(np.cross(z, x))
    poses = np.zeros((num_cameras, 4, 4), dtype=np.float64)
    poses[:, :3, :] = np.stack([x, y, z, origin], axis=-1)
    poses[:, 3, 3] = 1.0
    if manual_rotation is not None:
        T = np.eye(4, dtype=np.float64)
        T[:3, :3] = np.load(manual_rotation) if isinstance(manual_rotation, str) else manual_rotation
        poses = T @ poses
    return poses
