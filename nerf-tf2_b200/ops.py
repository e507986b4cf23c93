"""
PSNR metric and functions of the reference's `nerf/core/ops.py` (:187-257).

`PSNRMetric` divides the summed squared error by the number of RAYS, not rays*3
(core/ops.py:217,227), so it reads 10*log10(3) = 4.77 dB lower than `psnr_metric` /
`psnr_metric_numpy` on the same data. Both definitions are reproduced.
The metric state lives on the device as two floats [sq_error, count] that the loss kernel
accumulates into (nerfb200_mse_loss_grad), so `update_state` costs no extra pass.
"""
import math

import numpy as np
import torch


class PSNRMetric:
    """ops.PSNRMetric (core/ops.py:187-238)."""

    def __init__(self, name="psnr_metric", device=None):
        self.name = name
        self.state = None
        self._device = device

    def _ensure(self, device):
        if self.state is None:
            self.state = torch.zeros(2, device=device, dtype=torch.float32)

    def update_state(self, y_true, y_pred, sample_weight=None):
        self._ensure(y_true.device)
        self.state[0] += torch.sum(torch.square(y_true - y_pred))
        self.state[1] += float(y_true.shape[0])

    def result(self):
        if self.state is None:
            return float("nan")
        sq, cnt = self.state.tolist()
        if cnt == 0 or sq <= 0:
            return float("inf") if cnt else float("nan")
        return -10.0 * (math.log(sq / cnt) / math.log(10.0))

    def reset_states(self):
        if self.state is not None:
            self.state.zero_()

    reset_state = reset_states


def psnr_metric(y_true, y_pred):
    """ops.psnr_metric (core/ops.py:240-247) on torch tensors."""
    mse = torch.mean(torch.square(y_true - y_pred))
    return (-10.0) * (torch.log(mse) / math.log(10.0))


def psnr_metric_numpy(y_true, y_pred):
    """ops.psnr_metric_numpy (core/ops.py:249-257)."""
    mse = np.mean(np.square(y_true - y_pred))
    return (-10.) * (np.log(mse) / np.log(10.))
