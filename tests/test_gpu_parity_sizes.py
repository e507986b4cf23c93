"""GPU parity at the BASELINE.json sizes, against the CPU oracle computed live (tens of seconds of CPU each):

  cfg1  one full 100x100 synthetic 360-degree view, 64+128 samples: rendered rgb / depth / acc per pixel, every
        tensor-core precision, in the form stated in oracle/tolerance.py -- p99 AND a bounded maximum over every
        pixel, except the counted (<= 0.05 %) rays whose last-sample alpha sits on the other side of the
        reference's 0/1 jump (utils/ray_utils.py:459-468) -- and a PSNR against a structured ground truth;
  cfg3  one 4096-ray coarse+fine training step: loss, gradients against the ORACLE's autograd gradients, parameters
        after Adam;
  a13   NeRF.test_step's metric against the oracle's PSNRMetric (core/model.py:182-223).

Stated tolerances (oracle/tolerance.py; measured values in profiles/r2*_parity_*.json; W3 units for depth):

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

from oracle.tolerance import (MAX_FLIP_RATE, MAX_FLIP_RATE_COARSE_SAMPLING, RENDER_TOL, check_render, last_alpha_flips)


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
    v = cfg1["view"]
    nerf = make_nerf(cfg1["weights"], precision)
    oc, of = nerf.render_rays(dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]), u_fine=dev(cfg1["u_fine"]),
                              need_weights=True)
    npd = lambda d: {k: host(x) for k, x in d.items()}
    meas = check_render(precision, npd(oc), npd(of), cfg1["coarse"], cfg1["fine"])
    # PSNR against a STRUCTURED ground truth of the same scene: the oracle's fine image moved 10 % of the way towards its
    # coarse image (~29 dB from the fine image; uniform-random ground truth would sit at 8 dB and hide 1e-2 errors)
    fine, coarse = cfg1["fine"]["pred_rgb"], cfg1["coarse"]["pred_rgb"]
    gt = np.clip(fine + 0.1 * (coarse - fine), 0.0, 1.0).astype(F32)
    clip = lambda a: np.clip(a * 255.0, 0.0, 255.0) / 255.0
    p_gpu = rm.psnr_metric_numpy(gt, clip(host(of["pred_rgb"])))
    p_ref = rm.psnr_metric_numpy(gt, clip(fine))
    meas["psnr_vs_structured_gt"] = (float(p_gpu), float(p_ref))
    _record("cfg1_" + precision, meas)
    assert 22.0 <= p_ref <= 40.0, p_ref
    assert abs(p_gpu - p_ref) <= RENDER_TOL[precision]["psnr"], (precision, p_gpu, p_ref)


def test_split_last_sample_launch_is_what_bounds_the_maximum(cfg1):
    """Without the split launch bf16 leaves ~0.2 % of the rays of a full cfg1 view on the wrong side of the last-sample
    alpha jump (rgb errors of ~0.5); with it the count drops under the stated 0.05 %: the bounded maximum is a
    property of the kernel, not of a lucky view."""
    v = cfg1["view"]
    args = (dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]))
    npd = lambda d: {k: host(x) for k, x in d.items()}
    flips, errs = {}, {}
    for precise in (False, True):
        nerf = make_nerf(cfg1["weights"], "bf16", precise_last=precise)
        oc, of = nerf.render_rays(*args, u_fine=dev(cfg1["u_fine"]), need_weights=True)
        flips[precise] = int(last_alpha_flips(npd(oc), npd(of), cfg1["coarse"], cfg1["fine"]).sum())
        errs[precise] = np.abs(host(of["pred_rgb"]) - cfg1["fine"]["pred_rgb"]).max(axis=1)
    n = errs[True].shape[0]
    _record("cfg1_split_on_off", dict(n=n, flips_without=flips[False], flips_with=flips[True],
                                      max_err_without=float(errs[False].max()), max_err_with=float(errs[True].max())))
    assert flips[False] >= 10 and errs[False].max() > 0.1, (flips, errs[False].max())
    assert flips[True] <= 5e-4 * n and flips[True] * 4 <= flips[False], flips


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
    >= 0.9999 globally and >= 0.995 for each of the 48 tensors AGAINST THE ORACLE; the parameter update after Adam
    at cosine >= 0.995 with the oracle's update (measured: profiles/r2c_parity_train4096_bf16.json)."""
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
    assert abs(float(nerf.last_loss.item()) - info["loss"]) <= 3e-3 * info["loss"], (float(nerf.last_loss.item()), info["loss"])   # measured 9e-4
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
    assert worst[0] >= 0.995, worst            # measured 0.99905 (fine/dense_0/kernel: the high-frequency encoding columns)
    cos = float(torch.nn.functional.cosine_similarity(nerf.flat_grads.double(), flat_ref, dim=0))
    assert cos >= 0.9999, cos                  # measured 0.99998
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
    assert ucos >= 0.995, ucos                 # measured 0.9995
    # PSNRMetric of the step (fine prediction) against the oracle's
    mref = rm.PSNRMetric(); mref.update_state(gt, info["pred_rgb_f"])
    assert abs(float(logs["psnr_metric"]) - float(mref.result())) <= 2e-2


@pytest.mark.parametrize("cfg", [dict(Nc=128, Nf=256, lin=True, white=True, perturb=True),      # BASELINE cfg5's sample counts
                                 dict(Nc=32, Nf=64, lin=False, white=False, perturb=True),      # generic sampler / integrator paths
                                 dict(Nc=64, Nf=128, lin=False, white=False, perturb=False)])
def test_forward_other_configurations_vs_oracle(cfg):
    """The whole march at other settings of params.sampling / system.white_bg than the default render: 128+256 samples
    (cfg5), linear-in-depth spacing, black background, stratified perturbation with explicit uniforms -- fp32 MLP near-exact,
    bf16 within the stated tolerance, through the one-call entry point and step by step."""
    H, W = 20, 15
    n = H * W
    v = osc.synthetic_view(H, W, view=5)
    rng = np.random.default_rng(cfg["Nc"])
    uc = rng.random((n, cfg["Nc"]), dtype=F32) if cfg["perturb"] else None
    uf = rng.random((n, cfg["Nf"]), dtype=F32)
    w = om.init_weights(9)
    pc, pf, dbg = om.forward(w, v["rays_o"], v["rays_d"], v["near"], v["far"], cfg["Nc"], cfg["Nf"], lin_inv_depth=cfg["lin"],
                             perturb=cfg["perturb"], white_bg=cfg["white"], u_coarse=uc, u_fine=uf, return_debug=True)
    p = nb.make_params({"system": {"white_bg": cfg["white"]}}, N_coarse=cfg["Nc"], N_fine=cfg["Nf"], lin_inv_depth=cfg["lin"],
                       perturb=cfg["perturb"])
    args = (dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]))
    kw = dict(u_coarse=None if uc is None else dev(uc), u_fine=dev(uf))
    npd = lambda d: {k: host(x) for k, x in d.items()}
    f32 = nb.setup_model(p, precision="fp32"); f32.set_weights_from_dict(w)
    oc, of = f32.forward(*args, **kw)
    assert of["weights"].shape == (n, cfg["Nc"] + cfg["Nf"])
    for k, lim in (("pred_rgb", 5e-4), ("pred_depth", 1e-3), ("acc_map", 1e-3)):
        assert np.abs(host(oc[k]) - pc[k]).max() <= lim and np.abs(host(of[k]) - pf[k]).max() <= lim, (cfg, k)
    # the flip-rate bound of the stated tolerance is for the BASELINE sample counts; with 32+64 samples the last
    # hierarchical sample moves further for the same 16-bit error of the coarse pass (measured: 2 of 300 rays with bf16,
    # 0 with fp16, profiles/r2z_parity_other_cfg_32_64.json) -- there the bound is 1 %. In every configuration a
    # flipped ray must sit on the jump: the oracle's own last-sample pre-activation within ON_JUMP_MARGIN of zero.
    rate = MAX_FLIP_RATE if cfg["Nc"] >= 64 else MAX_FLIP_RATE_COARSE_SAMPLING
    pre_last = (dbg["sigma_pre_last_c"], dbg["sigma_pre_last_f"])
    meas = {}
    for prec in ("bf16", "fp16"):
        b16 = nb.setup_model(p, precision=prec); b16.set_weights_from_dict(w)
        for fused in (True, False):
            b16.fused_forward = fused
            oc, of = b16.forward(*args, **kw)
            meas[f"{prec}_{'one_call' if fused else 'step_by_step'}"] = check_render(prec, npd(oc), npd(of), pc, pf, max_flip_rate=rate,
                                                                                     pre_last=pre_last)
    _record(f"other_cfg_{cfg['Nc']}_{cfg['Nf']}", meas)
