"""
Configuration object with the attribute layout of the reference's python-box `params`
(utils/params_utils.py:4-13; keys and defaults from params/config.yaml). Nodes allow both
`params.data.batch_size` and `params.custom_dataset["train"]` access, like a Box.
`load_params` accepts a YAML path or a dict.
"""
import copy

DEFAULTS = {
    "system": {"tf_seed": 11, "white_bg": False, "run_eagerly": False, "log_images": False,
               "steps_per_epoch": 32, "validation_freq": 400, "tensorboard_dir": "./logs", "initial_epoch": 0,
               "dataset_type": "CustomDataset"},
    "eval": {"save_dir": "./output/eval"},
    "render": {"radius": 4.0, "inclination": 30.0, "num_cameras": 30, "img_size": [800, 800],
               "camera_model_name": "SIMPLE_PINHOLE", "camera_model_params": [1111.111, 400.0, 400.0],
               "bounds": None, "manual_rotation": None, "save_dir": "./output/render"},
    "model": {"save": {"save_dir": "./save_dir/models", "save_optimizer_state": True},
              "load": {"load_dir": "", "load_tag": "", "set_weights": False, "skip_optimizer": False}},
    "data": {"reconfig": {"save_dir": None, "load_dir": None}, "scale_imgs": None,
             "scene_scale_mul": 0.85, "scene_scale_add": 0.0, "batch_size": 4096, "dataset_mode": "sample",
             "sample_mode": {"shuffle_buffer_size": 20, "prefetch_buffer_size": 20, "repeat_count": 5000},
             "iterate_mode": {"repeat_count": 36, "train_shuffle": {"enable": True, "seed": 35},
                              "advance_train_tf_dataset": {"enable": False, "skip_count": 0}}},
    "blender_dataset": {"base_dir": "../lego", "shuffle": {"enable": ["train"], "seed": 83},
                        "val": {"num": 3, "frac": None}, "test": {"num": 3, "frac": None}},
    "custom_dataset": {"shuffle": {"enable": ["train"], "seed": 83},
                       "train": {"img_root_dir": "../data/train", "pose_info_path": "../data/train_pose_info.csv"},
                       "val": {"img_root_dir": "../data/val", "pose_info_path": "../data/val_pose_info.csv",
                               "num": 3, "frac": None},
                       "test": {"img_root_dir": "../data/test", "pose_info_path": "../data/test_pose_info.csv",
                                "num": 3, "frac": None}},
    "preprocessing": {"origin_method": "min_dist_solve", "bounds_method": "include_corners",
                      "basis_method": "compute", "manual_rotation": None},
    "sampling": {"N_coarse": 64, "N_fine": 128, "perturb": True, "lin_inv_depth": True},
}


class Node(dict):
    """dict with attribute access (the subset of python-box the reference relies on)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self[key] = value


def _to_node(d):
    if isinstance(d, dict):
        return Node({k: _to_node(v) for k, v in d.items()})
    return d


def _merge(base, over):
    out = copy.deepcopy(base)
    for k, v in (over or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def make_params(overrides=None, **sampling):
    """make_params({"system": {"white_bg": True}}, N_fine=256) -> attribute-access params."""
    d = _merge(DEFAULTS, overrides)
    d["sampling"].update(sampling)
    return _to_node(d)


def load_params(path_or_dict):
    """utils/params_utils.load_params: YAML file (or a dict) -> attribute-access object."""
    if isinstance(path_or_dict, dict):
        return make_params(path_or_dict)
    import yaml
    with open(path_or_dict, "r") as f:
        return make_params(yaml.safe_load(f))
