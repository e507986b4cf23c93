"""
TEST INFRASTRUCTURE ONLY -- writes tiny, deterministic datasets to disk in the two on-disk formats
the reference loads (core/datasets.py): a NeRF-synthetic "blender" scene (transforms_{split}.json +
{split}/r_{i}.png RGBA) and a custom scene in the "Pose Info Format" CSV. Used by
oracle/gen_golden_datasets.py (which runs the REFERENCE loaders over them) and by the loader tests
(which run this repo's loaders over byte-identical files). Nothing in the product imports this.
"""
import json
import os

import numpy as np

H, W = 12, 16
N_IMGS = {"train": 6, "val": 4, "test": 5}


def _poses_opengl(n, rng):
    """Cameras on a perturbed sphere of radius ~4 looking roughly at the origin, OpenGL convention
    (camera looks down -z, +y up), [n,4,4] float64."""
    out = []
    for _ in range(n):
        c = rng.normal(size=3)
        c = 4.0 * c / np.linalg.norm(c) * (1.0 + 0.05 * rng.normal())
        c[2] = abs(c[2]) + 0.3
        target = 0.15 * rng.normal(size=3)
        back = c - target
        back /= np.linalg.norm(back)
        right = np.cross([0.0, 0.0, 1.0], back)
        right /= np.linalg.norm(right)
        up = np.cross(back, right)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, back, c
        out.append(m)
    return np.array(out)


def _image_rgba(rng):
    img = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    img[:3, :, 3] = 0          # fully transparent rows
    img[3:6, :, 3] = 255       # fully opaque rows
    return img


def write_blender_scene(root, seed=7):
    """-> dict split -> list of (frame name, opengl pose). File numbering is deliberately not in
    lexicographic order (r_10 after r_9) to exercise the numeric sort."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    for split, n in N_IMGS.items():
        os.makedirs(os.path.join(root, split), exist_ok=True)
        poses = _poses_opengl(n, rng)
        numbers = list(range(n)) if split != "train" else [0, 2, 9, 10, 11, 1][:n]
        frames = []
        for k, num in enumerate(numbers):
            Image.fromarray(_image_rgba(rng), "RGBA").save(os.path.join(root, split, f"r_{num}.png"))
            frames.append({"file_path": f"./{split}/r_{num}", "rotation": 0.0123, "transform_matrix": poses[k].tolist()})
        # frames listed in reverse to make the loader's ordering matter
        meta = {"camera_angle_x": 0.6911112070083618, "frames": frames[::-1]}
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump(meta, f)
    return root


CAMERA_MODELS = [
    ("SIMPLE_PINHOLE", [20.5, 8.0, 6.0]),
    ("PINHOLE", [20.5, 21.0, 8.0, 6.0]),
    ("SIMPLE_RADIAL", [19.0, 8.25, 5.75, 0.01]),
    ("RADIAL", [19.0, 8.0, 6.0, 0.01, -0.002]),
    ("OPENCV", [22.0, 21.5, 7.5, 6.5, 0.01, 0.0, 0.001, 0.0]),
    ("FULL_OPENCV", [22.0, 21.5, 7.5, 6.5, 0.01, 0.0, 0.001, 0.0, 0.0, 0.0, 0.0, 0.0]),
]


def write_custom_scene(root, seed=11):
    """Pose Info Format: one CSV per split with image_name, camera_model, camera_params (a YAML list),
    pose (12 numbers, row-major 3x4 camera->world, Classic-CV) and near/far; opaque RGB PNGs."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    for split, n in N_IMGS.items():
        img_dir = os.path.join(root, split)
        os.makedirs(img_dir, exist_ok=True)
        poses = _poses_opengl(n, rng) @ flip
        rows = ["image_name,camera_model,camera_params,pose,near,far"]
        for k in range(n):
            name = f"img_{k:03d}.png"
            Image.fromarray(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8), "RGB").save(os.path.join(img_dir, name))
            model, mp = CAMERA_MODELS[k % len(CAMERA_MODELS)]
            pose12 = ", ".join(repr(float(x)) for x in poses[k][:3].reshape(-1))
            near, far = 1.5 + 0.1 * k, 6.0 + 0.25 * k
            rows.append(f'{name},{model},"[{", ".join(repr(float(x)) for x in mp)}]","[{pose12}]",{near!r},{far!r}')
        with open(os.path.join(root, f"{split}_pose_info.csv"), "w") as f:
            f.write("\n".join(rows) + "\n")
    return root


def config_overrides(kind, root, save_dir, dataset_mode="iterate", white_bg=True, scale_imgs=None):
    """The config.yaml keys a loader reads, as a plain dict."""
    cfg = {
        "system": {"white_bg": bool(white_bg and kind == "BlenderDataset"), "dataset_type": kind, "tf_seed": 11},
        "data": {"reconfig": {"save_dir": save_dir, "load_dir": save_dir}, "scale_imgs": scale_imgs,
                 "scene_scale_mul": 0.85, "scene_scale_add": 0.0, "batch_size": 50, "dataset_mode": dataset_mode,
                 "sample_mode": {"shuffle_buffer_size": 20, "prefetch_buffer_size": 20, "repeat_count": 3},
                 "iterate_mode": {"repeat_count": 2, "train_shuffle": {"enable": True, "seed": 35},
                                  "advance_train_tf_dataset": {"enable": False, "skip_count": 0}}},
        "blender_dataset": {"base_dir": root, "shuffle": {"enable": ["train"], "seed": 83},
                            "val": {"num": 3, "frac": None}, "test": {"num": None, "frac": 0.5}},
        "custom_dataset": {"shuffle": {"enable": ["train"], "seed": 83},
                           "train": {"img_root_dir": os.path.join(root, "train"),
                                     "pose_info_path": os.path.join(root, "train_pose_info.csv")},
                           "val": {"img_root_dir": os.path.join(root, "val"),
                                   "pose_info_path": os.path.join(root, "val_pose_info.csv"), "num": 3, "frac": None},
                           "test": {"img_root_dir": os.path.join(root, "test"),
                                    "pose_info_path": os.path.join(root, "test_pose_info.csv"), "num": None, "frac": 0.5}},
        "preprocessing": {"origin_method": "min_dist_solve", "bounds_method": "include_corners",
                          "basis_method": "compute", "manual_rotation": None},
    }
    return cfg
