"""
The training script's flow as a function (main/train.py:20-66): datasets from the configured loader, the number of
epochs from the dataset mode's repeat count, the model with its CustomSaver callback, `nerf.fit`.
"""
from . import datasets as _datasets
from .model import setup_model_and_callbacks


def plan_epochs(params, num_imgs, img_HW):
    """(total_steps, steps_per_epoch, num_epochs) exactly as main/train.py:32-47 computes them."""
    steps_per_epoch = params.system.steps_per_epoch
    if params.data.dataset_mode == "iterate":
        repeat = params.data.iterate_mode.repeat_count
        total_steps = int(img_HW[0] * img_HW[1] * num_imgs["train"] * repeat / params.data.batch_size)
    elif params.data.dataset_mode == "sample":
        total_steps = int(num_imgs["train"] * params.data.sample_mode.repeat_count)
    else:
        raise ValueError(f"Invalid dataset mode: {params.data.dataset_mode}")
    return total_steps, steps_per_epoch, int(total_steps / steps_per_epoch)


def launch(params, **model_kw):
    """Train as `python -m nerf.main.train --config ...` would; returns (nerf, history)."""
    tf_datasets, num_imgs, img_HW = _datasets.get_tf_datasets_and_metadata_for_splits(params)
    _, steps_per_epoch, num_epochs = plan_epochs(params, num_imgs, img_HW)
    nerf, callbacks = setup_model_and_callbacks(params, num_imgs, img_HW, **model_kw)
    hist = nerf.fit(x=tf_datasets["train"], epochs=num_epochs, validation_data=tf_datasets["val"],
                    validation_freq=params.system.validation_freq, callbacks=callbacks,
                    steps_per_epoch=steps_per_epoch, initial_epoch=params.system.initial_epoch)
    return nerf, hist
