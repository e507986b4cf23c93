"""
Checkpoint / resume: the `CustomSaver` callback (core/ops.py:88-185) and
`NeRF.set_everything` (core/model.py:239-287).

File set per save, named like the reference (`{epoch:06d}_{val_psnr:.2f}_...`):
  *_optimizer.npz   identical format: one array per optimiser variable keyed by its name plus a `names`
                    array preserving order ([iter, m x48, v x48], core/ops.py:110-120,146-149)
  *_logs.npz        the collected logs (core/ops.py:122-127)
  *_coarse.npz / *_fine.npz   the 24 variables per sub-model keyed by name plus `names` (default), or
  *_coarse.h5  / *_fine.h5    Keras `save_weights` files (core/ops.py:142-143) with `weights_format="h5"`.
                    There is no HDF5 library in this image; `.h5` goes through h5lite.py, whose READER is
                    checked against a real libhdf5 file and whose WRITER is round-trip checked only --
                    hence `.npz` stays the default for files this repo writes. `set_everything` reads
                    whichever exists, `.h5` (the reference's name) first.
"""
import os
from collections import defaultdict
from copy import deepcopy

import numpy as np


def optimizer_variable_names(nerf):
    names = ["Adam/iter:0"]
    for slot in ("m", "v"):
        names += [f"Adam/{v.name}/{slot}:0" for v in nerf.trainable_variables]
    return names


def keras_layers(model_name):
    """`model.layers` of get_coarse_or_fine_model(model_name) in Keras' order (see oracle/model.py for the
    derivation): every layer appears in `layer_names`, only the Dense ones carry weights."""
    order = ["xyz", "enc_xyz"] + [f"dense_{i}" for i in range(5)] + ["concat_1", "dense_5", "dense_6", "dense_7",
             "rays_d", "dense_8", "enc_rays_d", "concat_2", "dense_9", "rgb", "sigma"]
    return [f"{model_name}/{n}" for n in order]


def save_weights_h5(path, sub_model):
    """Model.save_weights(path.h5) for one sub-model (layer groups, `weight_names`, `<var>:0` datasets)."""
    from . import h5lite
    by_layer = {}
    for v in sub_model.trainable_variables:
        by_layer.setdefault(v.name.rsplit("/", 1)[0], []).append((v.name + ":0", v.numpy()))
    h5lite.save_keras_weights(path, [(ln, by_layer.get(ln, [])) for ln in keras_layers(sub_model.name)])


def load_weights(path, sub_model):
    """Model.load_weights for one sub-model: `.h5` (topological order, like Keras' by_name=False: the
    file's weighted layers must match the model's in number, order and shapes) or this repo's `.npz`."""
    want = [v for v in sub_model.trainable_variables]
    if path.endswith(".h5") or path.endswith(".hdf5"):
        from . import h5lite
        got = h5lite.load_keras_weights(path)
        assert len(got) == len(want), f"{path}: {len(got)} weight tensors, the model has {len(want)}"
        arrays = []
        for (name, arr), v in zip(got, want):
            assert tuple(arr.shape) == v.shape, f"{path}: {name} has shape {arr.shape}, expected {v.shape} ({v.name})"
            arrays.append(arr)
    else:
        data = np.load(path)
        names = [str(n) for n in data["names"]]
        assert names == [v.name for v in want], f"{path}: variable names/order differ"
        arrays = [data[n] for n in names]
    sub_model.set_weights(arrays)


def save_weights(path, variables):
    """CustomSaver._save_weights (core/ops.py:110-120)."""
    names = [v.name if hasattr(v, "name") else str(v[0]) for v in variables]
    values = [v.numpy() if hasattr(v, "numpy") else np.asarray(v[1]) for v in variables]
    items = {k: val for k, val in zip(names, values)}
    items["names"] = np.array(names)
    np.savez(path, **items)


def save_optimizer(path, nerf):
    names = optimizer_variable_names(nerf)
    values = nerf.optimizer.variables()
    items = {k: np.asarray(v) for k, v in zip(names, values)}
    items["names"] = np.array(names)
    np.savez(path, **items)


class CustomSaver:
    """ops.CustomSaver: saves sub-model weights, optimiser state and logs after each validation run."""

    def __init__(self, save_dir=None, save_best_only=False, save_optimizer_state=True, weights_format="npz", params=None):
        """`CustomSaver(params=params, save_best_only=False)` as in the reference (core/ops.py:98-108: the directory and
        the optimizer switch come from params.model.save), or the directory given directly."""
        if params is None and hasattr(save_dir, "model"):        # the reference passes params first
            params, save_dir = save_dir, None
        if params is not None:
            save_dir = params.model.save.save_dir if save_dir is None else save_dir
            save_optimizer_state = params.model.save.save_optimizer_state
        assert save_dir is not None, "CustomSaver needs params or a save directory"
        assert weights_format in ("npz", "h5")
        self.params = params
        self.weights_format = weights_format
        self.root = save_dir
        self.best_score = -1
        self.collected_logs = defaultdict(list)
        self.save_best_only = save_best_only
        self.save_opt_state = save_optimizer_state
        os.makedirs(self.root, exist_ok=True)
        self.model = None

    def set_model(self, model):
        self.model = model

    def _save_everything(self, epoch, val_psnr_score):
        name = f"{epoch:06d}_{val_psnr_score:.2f}"
        for sub in (self.model.coarse_model, self.model.fine_model):
            if self.weights_format == "h5":
                save_weights_h5(os.path.join(self.root, f"{name}_{sub.name}.h5"), sub)
            else:
                save_weights(os.path.join(self.root, f"{name}_{sub.name}.npz"), sub.trainable_variables)
        np.savez(os.path.join(self.root, f"{name}_logs.npz"), **deepcopy(dict(self.collected_logs)))
        if self.save_opt_state:
            save_optimizer(os.path.join(self.root, f"{name}_optimizer.npz"), self.model)
        return name

    def on_epoch_end(self, epoch, logs):
        self.collected_logs["train_epoch_idxs"].append(epoch)
        self.collected_logs["train_psnr_metric"].append(logs["psnr_metric"])
        if "val_psnr_metric" not in logs:
            return
        val = logs["val_psnr_metric"]
        self.collected_logs["val_epoch_idxs"].append(epoch)
        self.collected_logs["val_psnr_metric"].append(val)
        if self.save_best_only and not val > self.best_score:
            return
        self._save_everything(epoch, val)
        self.best_score = max(self.best_score, val)


def set_everything(nerf, load_dir, load_tag, skip_optimizer=False):
    """NeRF.set_everything (core/model.py:239-287): restores sub-model weights and the optimiser state,
    asserting that the saved optimiser variable names match the current ones in order."""
    if nerf.optimizer is None:
        nerf.compile()
    if not skip_optimizer:
        data = np.load(os.path.join(load_dir, f"{load_tag}_optimizer.npz"))
        saved = [str(n) for n in data["names"]]
        current = optimizer_variable_names(nerf)
        assert current == saved, ("The optimizer state cannot be loaded since the current variable names "
                                  "and saved variable names are different.")
        nerf.optimizer.set_weights([data[n] for n in saved])
    for sub, tag in ((nerf.coarse_model, "coarse"), (nerf.fine_model, "fine")):
        h5 = os.path.join(load_dir, f"{load_tag}_{tag}.h5")
        load_weights(h5 if os.path.exists(h5) else os.path.join(load_dir, f"{load_tag}_{tag}.npz"), sub)
