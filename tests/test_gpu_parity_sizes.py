"""GPU parity at the BASELINE.json sizes, against the CPU oracle computed live (tens of seconds of CPU each):

  cfg1  one full 100x100 synthetic 360-degree view, 64+128 samples: rendered rgb / depth / acc per pixel, every
        tensor-core precision, with BOUNDED maxima (no outlier allowance: the split last-sample launch removes the
        sigma_last sign-flip class, utils/ray_utils.py:459-468) and a PSNR against a structured ground truth;
  cfg3  one 4096-ray coarse+fine training step: loss, gradients against the ORACLE's autograd gradients, parameters
        after Adam;
  a13   NeRF.test_step's metric against the oracle's PSNRMetric (core/model.py:182-223).

Stated tolerances (measured values in profiles/r2a_parity_diag.json; W3 units for depth, near/far 0.425/1.275):

  precision  rgb p99 / max     depth p99 / max   acc p99 / max     PSNR vs structured GT
  bf16       4e-3 / 1.5e-2     8e-3 / 4e-2       5e-3 / 2e-2       0.05 dB
  fp16,tf32  2e-3 / 1.2e-2     3e-3 / 3e-2       2e-3 / 1.5e-2     0.02 dB
"""
import numpy as np
import pytest
import torch

import nerf_tf2_b200 as nb
from oracle import model as om, ray_march as rm, scene as osc

pytestmark = pytest.mark.gpu
F32 = np.float32

TOL = {"bf16": dict(rgb=(4e-3, 1.5e-2), depth=(8e-3, 4e-2), acc=(5e-3, 2e-2), psnr=0.05),
       "fp16": dict(rgb=(2e-3, 1.2e-2), depth=(3e-3, 3e-2), acc=(2e-3, 1.5e-2), psnr=0.02),
       "tf32": dict(rgb=(2e-3, 1.2e-2), depth=(3e-3, 3e-2), acc=(2e-3, 1.5e-2), psnr=0.02)}


def _record(name, d):
    """Measured values behind the stated tolerances, for profiles/ (written when gpurun_out/ exists)."""
    import json, os
    from conftest import ROOT
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"parity_{name}.json"), "w") as f:
            json.dump(d, f, indent=1)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def make_nerf(weights, precision, white_bg=True, perturb=False, **kw):
    p = nb.make_params({"system": {"white_bg": white_bg}}, perturb=perturb)
    nerf = nb.setup_model(p, precision=precision, **kw)
    nerf.set_weights_from_dict(weights)
    return nerf


@pytest.fixture(scope="module")
def cfg1():
    """BASELINE.json configs[0]: random-init coarse+fine, 64+128 samples, one 100x100 view; fp32 oracle render."""
    H = W = 100
    v = osc.synthetic_view(H, W, view=1)
    uf = np.random.default_rng(11).random((H * W, 128), dtype=F32)
    w = om.init_weights(7)
    pc, pf = om.forward(w, v["rays_o"], v["rays_d"], v["near"], v["far"], u_fine=uf, perturb=False, white_bg=True)
    return dict(H=H, W=W, view=v, u_fine=uf, weights=w, coarse=pc, fine=pf)


@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
def test_render_cfg1_full_view_vs_oracle(cfg1, precision):
    v, tol = cfg1["view"], TOL[precision]
    nerf = make_nerf(cfg1["weights"], precision)
    oc, of = nerf.render_rays(dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]), u_fine=dev(cfg1["u_fine"]),
                              need_weights=True)
    n = cfg1["H"] * cfg1["W"]
    meas = {}
    for name, out, ref in (("coarse", oc, cfg1["coarse"]), ("fine", of, cfg1["fine"])):
        for key, tk in (("pred_rgb", "rgb"), ("pred_depth", "depth"), ("acc_map", "acc")):
            e = np.abs(host(out[key]).reshape(n, -1) - ref[key].reshape(n, -1)).max(axis=1)
            p99, mx = float(np.percentile(e, 99)), float(e.max())
            meas[f"{name}_{key}"] = (p99, mx)
            assert p99 <= tol[tk][0] and mx <= tol[tk][1], (precision, name, key, p99, mx)
        # no outlier class is left: not one pixel beyond the stated maximum (the unbounded 1 % allowance is gone)
    # PSNR against a STRUCTURED ground truth of the same scene (the oracle's coarse image: ~25-35 dB from the fine one)
    gt = np.clip(cfg1["coarse"]["pred_rgb"], 0.0, 1.0).astype(F32)
    clip = lambda a: np.clip(a * 255.0, 0.0, 255.0) / 255.0
    p_gpu = rm.psnr_metric_numpy(gt, clip(host(of["pred_rgb"])))
    p_ref = rm.psnr_metric_numpy(gt, clip(cfg1["fine"]["pred_rgb"]))
    assert 15.0 <= p_ref <= 60.0, p_ref
    meas["psnr_vs_structured_gt"] = (float(p_gpu), float(p_ref))
    _record("cfg1_" + precision, meas)
    assert abs(p_gpu - p_ref) <= tol["psnr"], (precision, p_gpu, p_ref)


def test_split_last_sample_launch_is_what_bounds_the_maximum(cfg1):
    """Without the split launch bf16 leaves sign-flip outliers of ~0.5 in rgb on a full cfg1 view; with it none:
    the bounded maxima above are a property of the kernel, not of a lucky view."""
    v = cfg1["view"]
    args = (dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]))
    errs = {}
    for precise in (False, True):
        nerf = make_nerf(cfg1["weights"], "bf16", precise_last=precise)
        _, of = nerf.render_rays(*args, u_fine=dev(cfg1["u_fine"]))
        errs[precise] = np.abs(host(of["pred_rgb"]) - cfg1["fine"]["pred_rgb"]).max(axis=1)
    assert errs[False].max() > 0.1 and (errs[False] > 5e-2).mean() > 5e-4, errs[False].max()
    assert errs[True].max() <= TOL["bf16"]["rgb"][1] and (errs[True] > 5e-2).mean() == 0.0


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_test_step_metric_vs_oracle(cfg1, precision):
    """a13: NeRF.test_step / evaluate (core/model.py:182-223) -- the PSNRMetric of the FINE prediction, accumulated
    over batches, against the oracle's PSNRMetric on the oracle's fine prediction (same fixed uniforms)."""
    v = cfg1["view"]
    n = 2048
    rng = np.random.default_rng(5)
    gt = np.clip(cfg1["coarse"]["pred_rgb"][:n] + rng.normal(scale=0.02, size=(n, 3)), 0, 1).astype(F32)
    nerf = make_nerf(cfg1["weights"], precision)
    ref = rm.PSNRMetric()
    logs = None
    for s0 in (0, 1024):
        sl = slice(s0, s0 + 1024)
        batch = ((v["rays_o"][sl], v["rays_d"][sl], v["near"][sl], v["far"][sl]), (gt[sl],))
        logs = nerf.test_step(batch, u_fine=dev(cfg1["u_fine"][sl]))
        ref.update_state(gt[sl], cfg1["fine"]["pred_rgb"][sl])
    tol = 1e-3 if precision == "fp32" else 2e-2
    assert abs(float(logs["psnr_metric"]) - float(ref.result())) <= tol, (float(logs["psnr_metric"]), float(ref.result()))
    assert abs(nerf.metrics[0].result() - float(ref.result())) <= tol
    # evaluate() = reset + test_step over the dataset + result (perturbation off, in-kernel fine uniforms: looser)
    ds = nb.RayDataset.from_tensor_slices(((v["rays_o"][:n], v["rays_d"][:n], v["near"][:n], v["far"][:n]), (gt,))).batch(1024)
    assert abs(nerf.evaluate(ds) - float(ref.result())) <= 0.3


@pytest.mark.parametrize("precision", ["bf16"])
def test_train_step_cfg3_4096_rays_vs_oracle(precision):
    """BASELINE.json configs[2] at full size: 4096 rays of an 800x800 view, coarse+fine forward/backward + Adam, against
    the oracle's fp32 autograd gradients (18 s, 14 GB of CPU). Stated tolerance: loss 3e-3 relative; gradient cosine
    >= 0.9995 globally and >= 0.97 for each of the 48 tensors AGAINST THE ORACLE; the parameter update after Adam
    at cosine >= 0.90 with the oracle's update."""
    v = osc.synthetic_view(800, 800, view=0)
    rng = np.random.default_rng(3)
    sel = rng.choice(640000, size=4096, replace=False)
    uf = rng.random((4096, 128), dtype=F32)
    gt = rng.random((4096, 3), dtype=F32)
    ro, rd, near, far = (np.ascontiguousarray(v[k][sel]) for k in ("rays_o", "rays_d", "near", "far"))
    w0 = om.init_weights(7)
    w_ref = {k: a.copy() for k, a in w0.items()}
    info, g_ref = om.loss_and_grads(w_ref, ro, rd, near, far, gt, u_fine=uf)
    names = om.all_variable_names()

    nerf = make_nerf(w0, precision, train_precision=precision)
    before = nerf.flat_params.clone()
    logs = nerf.train_step(((ro, rd, near, far), (gt,)), u_fine=dev(uf))
    assert abs(float(nerf.last_loss.item()) - info["loss"]) <= 3e-3 * info["loss"], (float(nerf.last_loss.item()), info["loss"])
    var = {v_.name: v_ for v_ in nerf.trainable_variables}
    flat_ref = torch.zeros_like(nerf.flat_grads, dtype=torch.float64)
    worst = (1.0, None)
    for nme in names:
        vv = var[nme]
        a = nerf.flat_grads[vv._ofs:vv._ofs + vv._n].double().cpu()
        b = torch.from_numpy(g_ref[nme].reshape(-1)).double()
        flat_ref[vv._ofs:vv._ofs + vv._n] = b.cuda()
        c = float(torch.nn.functional.cosine_similarity(a, b, dim=0))
        worst = min(worst, (c, nme))
    assert worst[0] >= 0.97, worst
    cos = float(torch.nn.functional.cosine_similarity(nerf.flat_grads.double(), flat_ref, dim=0))
    assert cos >= 0.9995, cos
    # Adam (Keras OptimizerV2 form) on the oracle's gradients vs the fused kernel on the device's
    m = {k: np.zeros_like(a) for k, a in w_ref.items()}
    vv_ = {k: np.zeros_like(a) for k, a in w_ref.items()}
    om.adam_step(w_ref, g_ref, m, vv_, 0)
    upd_ref = torch.zeros_like(flat_ref)
    for nme in names:
        x = var[nme]
        upd_ref[x._ofs:x._ofs + x._n] = torch.from_numpy((w_ref[nme] - w0[nme]).reshape(-1)).double().cuda()
    # (the FIRST Adam step moves every parameter by ~lr * sign(g): the update compares gradient SIGNS element by element,
    # the harshest view of a reduced-precision gradient -- stated as the cosine between the two update vectors)
    upd = (nerf.flat_params - before).double()
    ucos = float(torch.nn.functional.cosine_similarity(upd, upd_ref, dim=0))
    _record("train4096_" + precision, dict(loss=float(nerf.last_loss.item()), loss_ref=info["loss"], grad_cos=cos,
                                           worst_tensor_cos=worst[0], worst_tensor=worst[1], update_cos=ucos))
    assert ucos >= 0.90, ucos
    # PSNRMetric of the step (fine prediction) against the oracle's
    mref = rm.PSNRMetric(); mref.update_state(gt, info["pred_rgb_f"])
    assert abs(float(logs["psnr_metric"]) - float(mref.result())) <= 2e-2
