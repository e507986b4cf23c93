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


class MetricValue:
    """A PSNRMetric result that stays on the device until it is used as a number (float(), formatting, comparisons,
    arithmetic, NumPy conversion all work and synchronise once)."""
    __slots__ = ("_snap", "_val")

    def __init__(self, snapshot):
        self._snap, self._val = snapshot, None

    def _get(self):
        if self._val is None:
            sq, cnt = self._snap.tolist()
            self._val, self._snap = PSNRMetric._psnr(sq, cnt), None
        return self._val

    def __float__(self): return self._get()
    def __array__(self, dtype=None, copy=None): return np.asarray(self._get(), dtype=dtype)
    def __repr__(self): return repr(self._get())
    def __str__(self): return str(self._get())
    def __format__(self, spec): return format(self._get(), spec)
    def __bool__(self): return bool(self._get())
    def __hash__(self): return hash(self._get())
    def __eq__(self, o): return self._get() == float(o)
    def __lt__(self, o): return self._get() < float(o)
    def __le__(self, o): return self._get() <= float(o)
    def __gt__(self, o): return self._get() > float(o)
    def __ge__(self, o): return self._get() >= float(o)
    def __neg__(self): return -self._get()
    def __abs__(self): return abs(self._get())
    def __round__(self, n=None): return round(self._get(), n)
    def __add__(self, o): return self._get() + float(o)
    def __radd__(self, o): return float(o) + self._get()
    def __sub__(self, o): return self._get() - float(o)
    def __rsub__(self, o): return float(o) - self._get()
    def __mul__(self, o): return self._get() * float(o)
    def __rmul__(self, o): return float(o) * self._get()
    def __truediv__(self, o): return self._get() / float(o)
    def __rtruediv__(self, o): return float(o) / self._get()


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
        """state += [sum((y_true - y_pred)^2), number of rays] (core/ops.py:204-220), accumulated by the same kernel
        that accumulates it during train_step (nerfb200_mse_loss_grad with no loss / gradient outputs)."""
        from ._lib import check, load, ptr, stream_ptr
        y_true = y_true.to(torch.float32).contiguous()
        y_pred = y_pred.to(torch.float32).contiguous()
        assert y_true.shape == y_pred.shape and y_true.dim() == 2 and y_true.shape[1] == 3
        self._ensure(y_true.device)
        B = int(y_true.shape[0])
        if not y_true.is_cuda:       # a metric object fed HOST arrays (like psnr_metric_numpy): host arithmetic
            self.state[0] += torch.sum(torch.square(y_true - y_pred))
            self.state[1] += float(B)
            return
        with torch.cuda.device(y_true.device):
            check(load().nerfb200_mse_loss_grad(B, B, ptr(y_pred), ptr(y_true), None, None, ptr(self.state), stream_ptr()),
                  "mse_loss_grad")

    @staticmethod
    def _psnr(sq, cnt):
        if cnt == 0 or sq <= 0:
            return float("inf") if cnt else float("nan")
        return -10.0 * (math.log(sq / cnt) / math.log(10.0))

    def result(self):
        if self.state is None:
            return float("nan")
        sq, cnt = self.state.tolist()
        return self._psnr(sq, cnt)

    def result_async(self):
        """The metric value as of now WITHOUT synchronising with the device: a tiny device-side snapshot of the state,
        turned into a Python float only when somebody looks at it. train_step / test_step return these (Keras hands
        back tensors from its step functions for the same reason): a training loop that does not read the per-step
        logs never stalls the launch queue."""
        if self.state is None:
            return float("nan")
        return MetricValue(self.state.clone())

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
