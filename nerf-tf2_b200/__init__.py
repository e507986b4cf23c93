"""
nerf-tf2_b200: a B200-native (sm_100a) implementation of the NeRF ray-march hot path of
thatbrguy/nerf-tf2 -- camera rays, stratified + hierarchical sampling, positional encoding,
the coarse/fine 8x256 MLPs and the volume-rendering integrator -- behind the reference's own
Python surface (`NeRF.fit/evaluate/predict`, the `ray_utils` functions and result dicts).

Hand-written CUDA kernels in csrc/ are reached through the C ABI of include/nerfb200.h via
ctypes. There is no CPU fallback, no Triton and no alternative backend.
"""
from . import _lib, checkpoint, data, datasets, dist, model, ops, params, pose_utils, ray_utils, render, scene, train  # noqa: F401
from ._lib import NerfB200Error  # noqa: F401
from .checkpoint import CustomSaver  # noqa: F401
from .data import RayDataset, SampleModeDataset, create_dataset_for_render  # noqa: F401
from .datasets import (BlenderDataset, CustomDataset, RayLevelData, SceneLevelData, get_data_and_metadata_for_splits,  # noqa: F401
                       get_dataset_obj, get_tf_datasets_and_metadata_for_splits)
from .model import NeRF, PositionalEncoder, get_coarse_or_fine_model, setup_model, setup_model_and_callbacks  # noqa: F401
from .ops import PSNRMetric, psnr_metric, psnr_metric_numpy  # noqa: F401
from .params import load_params, make_params  # noqa: F401

__version__ = "0.1.0"
