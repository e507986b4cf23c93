"""
Dataset loaders and the scene -> ray preparation around the hot path:
the reference's core/base_dataset.py (`Dataset`) and core/datasets.py (`BlenderDataset`,
`CustomDataset`, factories), with the same class/method names, config keys and return structure.

What is different, B200-first: scene-level data (images, poses, bounds, intrinsics) is parsed and
normalised on the host exactly as the reference does (it is O(#images) work), but everything
per-RAY happens on the device -- `process_data` runs the get_rays kernel per image and keeps the ray
tensors in HBM, the "tf.data" objects returned are `RayDataset` (device-resident, batched views) and
`SampleModeDataset` (Philox pixel draw + ray generation + colour gather in HBM, data.py), so no ray
ever crosses PCIe during training or evaluation.

File formats: Blender = `transforms_{split}.json` + `{split}/r_{i}.png` (core/datasets.py:16-311);
Custom = the "Pose Info Format" CSV with columns image_name, camera_model, camera_params, pose, near,
far (core/datasets.py:313-585). PNGs are read through PIL (the reference's imageio "PNG-PIL" plugin is
a PIL wrapper; imageio itself is not installed here) or cv2, as there.
"""
import json
import os
from collections import namedtuple

import numpy as np

from . import pose_utils

SceneLevelData = namedtuple("SceneLevelData", ("imgs", "poses", "bounds", "intrinsics"))   # base_dataset.py:15-17
RayLevelData = namedtuple("RayLevelData", ("rays_o", "rays_d", "near", "far", "rgb"))      # base_dataset.py:19-22

SPLITS = ("train", "test", "val")      # the reference's iteration order (base_dataset.py:36)


def _first_k(items, split_params):
    """`num` / `frac` selection of val/test entries (core/datasets.py:135-160, 437-462)."""
    frac, num = split_params.get("frac"), split_params.get("num")
    if frac is not None and num is not None:
        raise ValueError("Either frac could be not None, OR num could be not None.")
    if frac is not None:
        assert type(frac) is float and 0 < frac < 1
        k = int(len(items) * frac)
    elif num is not None:
        assert type(num) is int
        k = num
    else:
        return items
    assert k <= len(items)
    return items[:k]


class Dataset:
    """Validation, W1->W2->W3 reconfiguration, ray extraction and dataset construction shared by the
    loaders (core/base_dataset.py:24-861). Subclasses provide `get_data_and_metadata_for_splits`."""

    def __init__(self, params):
        self.params = params
        self.splits = list(SPLITS)
        mul, add = params.data.scene_scale_mul, params.data.scene_scale_add
        self._scale_mul = 1.0 if mul is None else mul
        self._scale_add = 0.0 if add is None else add
        assert params.data.dataset_mode in ("sample", "iterate")
        self.sample_mode_params = params.data.sample_mode
        self.iterate_mode_params = params.data.iterate_mode

    # ---- to be provided by the loaders (base_dataset.py:52-109)
    def get_data_and_metadata_for_splits(self):
        raise NotImplementedError

    def get_tf_datasets_and_metadata_for_splits(self):
        """(datasets, num_imgs, img_HW): the name is the reference's; the values are RayDataset /
        SampleModeDataset objects (core/datasets.py:258-311, 532-585)."""
        data_splits, num_imgs = self.get_data_and_metadata_for_splits()
        if self.params.data.dataset_mode == "iterate":
            prepared, img_HW = self.prepare_data_iterate_mode(data_splits)
            datasets = self.create_tf_datasets_iterate_mode(prepared)
        else:
            prepared, img_HW = self.prepare_data_sample_mode(data_splits)
            datasets = self.create_tf_datasets_sample_mode(prepared)
        return datasets, num_imgs, img_HW

    # ---- validation (base_dataset.py:111-172)
    @staticmethod
    def _validate_intrinsic_matrix(K):
        """Only K = [[fu,0,cu],[0,fv,cv],[0,0,1]] is supported."""
        assert K.shape == (3, 3)
        assert K[2, 2] == 1
        assert K[0, 1] == 0 and K[1, 0] == 0 and K[2, 0] == 0 and K[2, 1] == 0

    def _validate_all_splits(self, data_splits):
        """Every image of every split has the same H x W; every intrinsic matrix is valid."""
        sizes = set()
        for split in self.splits:
            d = data_splits[split]
            for img, K in zip(d.imgs, d.intrinsics):
                sizes.add(tuple(img.shape[:2]))
                self._validate_intrinsic_matrix(K)
        assert len(sizes) <= 1, f"images of different sizes in the dataset: {sorted(sizes)}"

    # ---- reconfig parameters on disk (base_dataset.py:174-202)
    def _save_reconfig_params(self, W1_to_W2_transform, adj_scale_factor):
        root = self.params.data.reconfig.save_dir
        os.makedirs(root, exist_ok=True)
        np.savez(os.path.join(root, "reconfig.npz"), W1_to_W2_transform=W1_to_W2_transform,
                 adj_scale_factor=adj_scale_factor)

    def load_reconfig_params(self):
        d = np.load(os.path.join(self.params.data.reconfig.load_dir, "reconfig.npz"))
        return d["W1_to_W2_transform"], d["adj_scale_factor"]

    # ---- W1 -> W2 -> W3 (base_dataset.py:204-407)
    def _reconfigure_imgs_and_intrinsics(self, data_splits):
        out = {}
        for split in self.splits:
            d = data_splits[split]
            imgs, Ks = pose_utils.scale_imgs_and_intrinsics(d.imgs, d.intrinsics, self.params.data.scale_imgs)
            out[split] = SceneLevelData(imgs, d.poses, d.bounds, Ks)
        return out

    def _all(self, data_splits, field):
        return np.concatenate([getattr(data_splits[s], field) for s in self.splits], axis=0)

    def _reconfigure_poses(self, data_splits):
        pp = self.params.preprocessing
        T = pose_utils.calculate_new_world_transform(self._all(data_splits, "poses"), origin_method=pp.origin_method,
                                                     basis_method=pp.basis_method, manual_rotation=pp.manual_rotation)
        out = {s: data_splits[s]._replace(poses=pose_utils.reconfigure_poses(data_splits[s].poses, T))
               for s in self.splits}
        return out, T

    def _reconfigure_scene_scale(self, data_splits):
        H, W = data_splits["train"].imgs[0].shape[:2]
        scale = pose_utils.calculate_scene_scale(self._all(data_splits, "poses"), self._all(data_splits, "bounds"),
                                                 bounds_method=self.params.preprocessing.bounds_method,
                                                 intrinsics=self._all(data_splits, "intrinsics"), height=H, width=W)
        adj = scale * self._scale_mul + self._scale_add
        out = {}
        for s in self.splits:
            poses, bounds = pose_utils.reconfigure_scene_scale(data_splits[s].poses, data_splits[s].bounds, adj)
            out[s] = data_splits[s]._replace(poses=poses, bounds=bounds)
        return out, adj

    def validate_and_reconfigure_data(self, data_splits):
        """validate -> optional image rescale -> W1->W2 -> W2->W3 -> save reconfig.npz (base_dataset.py:365-407)."""
        self._validate_all_splits(data_splits)
        step1 = self._reconfigure_imgs_and_intrinsics(data_splits)
        step2, T = self._reconfigure_poses(step1)
        out, adj = self._reconfigure_scene_scale(step2)
        if self.params.data.reconfig.save_dir is not None:
            self._save_reconfig_params(T, adj)
        return out

    # ---- scene level -> ray level (base_dataset.py:409-468), on the device
    def process_data(self, data, on_device=True):
        """All H*W rays of every image (get_rays kernel: fp64 maths, fp32 results), near/far broadcast,
        rgb/255. Returns RayLevelData of CUDA tensors (or NumPy arrays with on_device=False)."""
        import torch
        from . import ray_utils
        ro, rd, near, far, rgb = [], [], [], [], []
        c255 = torch.tensor(255.0, device=ray_utils._dev(None), dtype=torch.float32)
        for img, pose, bound, K in zip(data.imgs, data.poses, data.bounds, data.intrinsics):
            H, W = img.shape[:2]
            o, d = ray_utils.get_rays(H, W, K, pose)
            n = o.shape[0]
            ro.append(o)
            rd.append(d)
            near.append(torch.full((n, 1), float(np.float32(bound[0])), device=o.device, dtype=torch.float32))
            far.append(torch.full((n, 1), float(np.float32(bound[1])), device=o.device, dtype=torch.float32))
            # IEEE division like the reference's `rgb / 255` (a tensor divisor: torch turns a scalar one into x * (1/255))
            rgb.append(torch.as_tensor(np.ascontiguousarray(img).reshape(-1, 3), device=o.device).to(torch.float32) / c255)
        out = RayLevelData(*(torch.cat(x, dim=0) for x in (ro, rd, near, far, rgb)))
        if not on_device:
            out = RayLevelData(*(t.cpu().numpy() for t in out))
        return out

    def prepare_data_iterate_mode(self, data_splits):
        """(RayLevelData per split, (H, W)) (base_dataset.py:470-505)."""
        reconf = self.validate_and_reconfigure_data(data_splits)
        return {s: self.process_data(reconf[s]) for s in self.splits}, reconf["train"].imgs[0].shape[:2]

    def prepare_data_sample_mode(self, data_splits):
        """train stays scene-level, val/test become ray-level (base_dataset.py:507-553)."""
        reconf = self.validate_and_reconfigure_data(data_splits)
        out = {s: (reconf[s] if s == "train" else self.process_data(reconf[s])) for s in self.splits}
        return out, out["train"].imgs[0].shape[:2]

    # ---- dataset objects (base_dataset.py:623-795)
    def _shuffle(self, rld):
        """One fixed permutation of all training rays (base_dataset.py:623-651)."""
        perm = np.random.default_rng(seed=self.iterate_mode_params.train_shuffle.seed).permutation(len(rld.rays_o))
        if not isinstance(rld.rays_o, np.ndarray):
            import torch
            perm = torch.from_numpy(perm).to(rld.rays_o.device)
        return RayLevelData(*(a[perm] for a in rld))

    @staticmethod
    def _separate(rld):
        return (rld.rays_o, rld.rays_d, rld.near, rld.far), (rld.rgb,)

    def _eval_dataset(self, rld):
        from .data import RayDataset
        return RayDataset.from_tensor_slices(self._separate(rld)).batch(self.params.data.batch_size, drop_remainder=False)

    def create_tf_datasets_iterate_mode(self, processed_splits):
        """train: (shuffled) rays in full batches, repeated `repeat_count` times; val/test: ragged last
        batch kept (base_dataset.py:663-725)."""
        from .data import RayDataset
        train = processed_splits["train"]
        if self.iterate_mode_params.train_shuffle.enable:
            train = self._shuffle(train)
        train_ds = RayDataset.from_tensor_slices(self._separate(train)).batch(self.params.data.batch_size,
                                                                              drop_remainder=True)
        train_ds = train_ds.repeat(count=self.iterate_mode_params.repeat_count)
        return {"train": train_ds, "test": self._eval_dataset(processed_splits["test"]),
                "val": self._eval_dataset(processed_splits["val"])}

    def create_tf_datasets_sample_mode(self, processed_splits):
        """train: one image per step, `batch_size` random pixels of it, drawn/generated/gathered on the
        device (base_dataset.py:555-621, 727-795)."""
        from .data import SampleModeDataset
        t = processed_splits["train"]
        train_ds = SampleModeDataset(t.imgs.astype(np.uint8), t.poses.astype(np.float32), t.bounds.astype(np.float32),
                                     t.intrinsics.astype(np.float32), batch_size=self.params.data.batch_size,
                                     seed=self.params.system.tf_seed, repeat_count=self.sample_mode_params.repeat_count)
        return {"train": train_ds, "test": self._eval_dataset(processed_splits["test"]),
                "val": self._eval_dataset(processed_splits["val"])}

    # ---- one view for rendering (base_dataset.py:797-861)
    def create_dataset_for_render(self, H, W, c2w, bounds, intrinsic, reconfig_poses):
        """`c2w` is camera->W1 (reconfig_poses=True) or camera->W2 (False); the saved reconfig.npz brings
        it to W3; rays are generated on the device and batched with drop_remainder=False."""
        from .data import create_dataset_for_render
        self._validate_intrinsic_matrix(K=intrinsic)
        T, adj = self.load_reconfig_params()
        pose = pose_utils.reconfigure_poses(c2w, T) if reconfig_poses else c2w
        pose, new_bounds = pose_utils.reconfigure_scene_scale(pose, bounds, adj)
        return create_dataset_for_render(H, W, pose, new_bounds, intrinsic, batch_size=self.params.data.batch_size)


class BlenderDataset(Dataset):
    """NeRF-synthetic ("blender") scenes (core/datasets.py:16-311)."""

    OPENGL_TO_CLASSIC_CV = np.diag([1.0, -1.0, -1.0, 1.0])      # flips the camera y and z axes (:69-82)

    def __init__(self, params):
        super().__init__(params)
        self.white_bg = params.system.white_bg
        self.blender_dataset_params = params.blender_dataset
        self.root = self.blender_dataset_params.base_dir
        self._configure_dataset()

    @staticmethod
    def _frame_name(frame):
        return frame["file_path"].split("/")[-1]

    def _configure_dataset(self):
        """Per split: image paths ordered by frame number, optionally shuffled (seeded), val/test cut to
        `num`/`frac`; frame name -> transform_matrix (core/datasets.py:32-67, 102-163)."""
        self.img_paths, self.metadata = {}, {}
        shuffle = self.blender_dataset_params.shuffle
        for split in SPLITS:
            with open(os.path.join(self.root, f"transforms_{split}.json"), "r") as f:
                meta = json.load(f)
            numbers = sorted(int(self._frame_name(fr).split("_")[-1].strip()) for fr in meta["frames"])
            paths = [os.path.join(self.root, split, f"r_{n}.png") for n in numbers]
            info = {"camera_angle_x": meta["camera_angle_x"]}
            for fr in meta["frames"]:
                info[self._frame_name(fr)] = {"rotation": fr.get("rotation"), "transform_matrix": fr["transform_matrix"]}
            self.metadata[split] = info
            if split in shuffle.enable:
                perm = np.random.default_rng(seed=shuffle.seed).permutation(len(paths))
                paths = np.array(paths)[perm].tolist()
            if split in ("test", "val"):
                paths = _first_k(paths, self.blender_dataset_params[split])
            self.img_paths[split] = paths

    @staticmethod
    def _opengl_to_classic_cv(pose):
        return np.asarray(pose) @ BlenderDataset.OPENGL_TO_CLASSIC_CV

    @staticmethod
    def _create_intrinsic_matrix(H, W, focal_length):
        return np.array([[focal_length, 0.0, W / 2], [0.0, focal_length, H / 2], [0.0, 0.0, 1.0]], dtype=np.float64)

    def _read_image(self, path):
        if self.white_bg:
            # RGBA composited on white in float32, truncated to uint8 (core/datasets.py:178-185)
            from PIL import Image
            rgba = np.array(Image.open(path)).astype(np.float32)
            alpha = rgba[..., 3] / 255.0
            return (rgba[..., :3] * alpha[..., None] + 255.0 * (1 - alpha[..., None])).astype(np.uint8)
        import cv2
        return cv2.cvtColor(cv2.imread(path, 1), cv2.COLOR_BGR2RGB)

    def _load_split(self, split):
        """Images, Classic-CV poses, bounds 2.0/6.0, one shared pinhole K from camera_angle_x
        (core/datasets.py:165-219); dtypes uint8 / float32 as there."""
        imgs, poses = [], []
        for path in self.img_paths[split]:
            name = os.path.basename(path).split(".")[0]
            imgs.append(self._read_image(path))
            poses.append(self._opengl_to_classic_cv(self.metadata[split][name]["transform_matrix"]))
        H, W = imgs[0].shape[:2]
        focal = 0.5 * W / np.tan(0.5 * float(self.metadata[split]["camera_angle_x"]))
        K = self._create_intrinsic_matrix(H, W, focal)
        n = len(imgs)
        return SceneLevelData(imgs=np.array(imgs).astype(np.uint8), poses=np.array(poses).astype(np.float32),
                              bounds=np.array([[2.0, 6.0]] * n).astype(np.float32),
                              intrinsics=np.array([K.copy() for _ in range(n)]).astype(np.float32))

    def get_data_and_metadata_for_splits(self):
        data = {s: self._load_split(s) for s in SPLITS}
        return data, {s: len(self.img_paths[s]) for s in SPLITS}


class CustomDataset(Dataset):
    """User datasets in the "Pose Info Format" CSV (core/datasets.py:313-585)."""

    SAME_FOCAL = frozenset(["SIMPLE_PINHOLE", "SIMPLE_RADIAL", "RADIAL"])      # f, cx, cy, ...
    DIFF_FOCAL = frozenset(["PINHOLE", "OPENCV", "FULL_OPENCV"])               # fx, fy, cx, cy, ...
    SUPPORTED_CAMERA_MODELS = SAME_FOCAL | DIFF_FOCAL

    def __init__(self, params):
        super().__init__(params)
        self.custom_dataset_params = params.custom_dataset
        self._configure_dataset()

    @classmethod
    def camera_model_params_to_intrinsics(cls, camera_model, model_params):
        """COLMAP camera model -> K; distortion terms are ignored (core/datasets.py:337-396)."""
        assert camera_model in cls.SUPPORTED_CAMERA_MODELS, f"Camera model {camera_model} is not supported."
        if camera_model in cls.SAME_FOCAL:
            fx, cx, cy = model_params[:3]
            fy = fx
        else:
            fx, fy, cx, cy = model_params[:4]
        return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float64)

    def _configure_dataset(self):
        import pandas as pd
        self.metadata = {}
        shuffle = self.custom_dataset_params.shuffle
        for split in SPLITS:
            sp = self.custom_dataset_params[split]
            df = pd.read_csv(sp.pose_info_path)
            if split in shuffle.enable:
                df = df.sample(frac=1, random_state=shuffle.seed).reset_index(drop=True)
            if split in ("test", "val"):
                k = len(_first_k(list(range(len(df))), sp))
                df = df.iloc[:k]
            self.metadata[split] = df

    def _load_split(self, split):
        import cv2
        import yaml
        sp = self.custom_dataset_params[split]
        imgs, poses, bounds, Ks = [], [], [], []
        for row in self.metadata[split].itertuples():
            imgs.append(cv2.cvtColor(cv2.imread(os.path.join(sp.img_root_dir, row.image_name), 1), cv2.COLOR_BGR2RGB))
            Ks.append(self.camera_model_params_to_intrinsics(row.camera_model, np.array(yaml.safe_load(row.camera_params))))
            poses.append(pose_utils.make_4x4(np.array(yaml.safe_load(row.pose)).reshape(3, 4)))
            bounds.append([row.near, row.far])
        return SceneLevelData(imgs=np.array(imgs).astype(np.uint8), poses=np.array(poses).astype(np.float32),
                              bounds=np.array(bounds).astype(np.float32), intrinsics=np.array(Ks).astype(np.float32))

    def get_data_and_metadata_for_splits(self):
        data = {s: self._load_split(s) for s in SPLITS}
        return data, {s: len(data[s].imgs) for s in SPLITS}


# ---------------------------------------------------------------- factories (core/datasets.py:587-755)
def get_dataset_obj(params):
    kind = params.system.dataset_type
    if kind == "BlenderDataset":
        return BlenderDataset(params=params)
    if kind == "CustomDataset":
        if params.system.white_bg:
            raise AssertionError("white_bg is only supported for BlenderDataset")
        return CustomDataset(params=params)
    raise ValueError(f"Invalid dataset type: {kind}")


def get_data_and_metadata_for_splits(params, return_dataset_obj=False):
    obj = get_dataset_obj(params)
    out = obj.get_data_and_metadata_for_splits()
    return (*out, obj) if return_dataset_obj else out


def get_tf_datasets_and_metadata_for_splits(params, return_dataset_obj=False):
    obj = get_dataset_obj(params)
    datasets, num_imgs, img_HW = obj.get_tf_datasets_and_metadata_for_splits()
    adv = params.data.iterate_mode.advance_train_tf_dataset
    if params.data.dataset_mode == "iterate" and adv.enable:
        datasets["train"] = datasets["train"].skip(adv.skip_count)
    return (datasets, num_imgs, img_HW, obj) if return_dataset_obj else (datasets, num_imgs, img_HW)
