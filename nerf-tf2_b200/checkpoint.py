"""
Checkpoint / resume: the `CustomSaver` callback (core/ops.py:88-185) and
`NeRF.set_everything` (core/model.py:239-287).

File set per save, named like the reference (`{epoch:06d}_{val_psnr:.2f}_...`):
  *_optimizer.npz   identical format: one array per optimiser variable keyed by its name plus a `names`
                    array preserving order ([iter, m x48, v x48], core/ops.py:110-120,146-149)
  *_logs.npz        the collected logs (core/ops.py:122-127)
  *_coarse.npz / *_fine.npz   the 24 variables per sub-model keyed by name plus `names`.
                    DEVIATION: the reference writes Keras `.h5` via save_weights (core/ops.py:142-143);
                    h5py/HDF5 is not available in this image, so the same arrays go into `.npz`.
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

    def __init__(self, save_dir, save_best_only=False, save_optimizer_state=True):
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
        save_weights(os.path.join(self.root, f"{name}_coarse.npz"), self.model.coarse_model.trainable_variables)
        save_weights(os.path.join(self.root, f"{name}_fine.npz"), self.model.fine_model.trainable_variables)
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
        data = np.load(os.path.join(load_dir, f"{load_tag}_{tag}.npz"))
        names = [str(n) for n in data["names"]]
        assert names == [v.name for v in sub.trainable_variables], f"{tag}: variable names/order differ"
        sub.set_weights([data[n] for n in names])
