"""GPU parity tests of the MLP paths and of the whole ray march / training step through the
reference-shaped Python surface, against the CPU oracle (fixtures in tests/golden/)."""
import numpy as np
import pytest
import torch

import nerf_tf2_b200 as nb
from nerf_tf2_b200 import _lib, ray_utils as ru
from oracle import model as om, ray_march as rm, scene as osc

pytestmark = pytest.mark.gpu
F32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def make_nerf(weights, precision, white_bg=True, perturb=False, **kw):
    p = nb.make_params({"system": {"white_bg": white_bg}}, perturb=perturb)
    nerf = nb.setup_model(p, precision=precision, **kw)
    nerf.set_weights_from_dict(weights)
    return nerf


# tolerances of the tensor-core paths, measured against the fp32 oracle (SURVEY.md App. E2 form:
# percentile + bounded outlier). rgb in [0,1]; sigma relative to max(1, sigma).
TC_TOL = {"bf16": dict(rgb_p99=6e-3, rgb_max=3e-2, sig_rel_p99=2e-2, sig_rel_max=1e-1),
          "fp16": dict(rgb_p99=1e-3, rgb_max=6e-3, sig_rel_p99=4e-3, sig_rel_max=3e-2),
          "tf32": dict(rgb_p99=1e-3, rgb_max=6e-3, sig_rel_p99=4e-3, sig_rel_max=3e-2)}


def test_mlp_fp32_path_matches_oracle(golden):
    g = golden["oracle_mlp"]
    w = om.init_weights(int(g["weights_seed"]), bias_scale=float(g["bias_scale"]))
    nerf = make_nerf(w, "fp32")
    for m, sub in (("coarse", nerf.coarse_model), ("fine", nerf.fine_model)):
        rgb, sigma = sub((dev(g["xyz"]), dev(g["dirs"])))
        assert sigma.shape == (g["xyz"].shape[0], 1)
        # fp32 FMA GEMMs vs MKL fp32 GEMMs: summation-order noise only (fp32-vs-fp64 is 2.4e-4, App. E2)
        assert np.abs(host(rgb) - g[f"{m}_rgb_f32"]).max() <= 5e-5
        assert np.allclose(host(sigma), g[f"{m}_sigma_f32"], rtol=2e-4, atol=2e-4)
        assert np.abs(host(rgb) - g[f"{m}_rgb_f64"]).max() <= 5e-4


@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
def test_mlp_tensor_core_path_matches_oracle(golden, precision):
    g = golden["oracle_mlp"]
    w = om.init_weights(int(g["weights_seed"]), bias_scale=float(g["bias_scale"]))
    nerf = make_nerf(w, precision)
    tol = TC_TOL[precision]
    for m, sub in (("coarse", nerf.coarse_model), ("fine", nerf.fine_model)):
        rgb, sigma = sub((dev(g["xyz"]), dev(g["dirs"])))
        e = np.abs(host(rgb) - g[f"{m}_rgb_f64"])
        assert np.percentile(e, 99) <= tol["rgb_p99"] and e.max() <= tol["rgb_max"], (np.percentile(e, 99), e.max())
        ref = g[f"{m}_sigma_f64"]
        es = np.abs(host(sigma) - ref) / np.maximum(1.0, np.abs(ref))
        assert np.percentile(es, 99) <= tol["sig_rel_p99"] and es.max() <= tol["sig_rel_max"], (np.percentile(es, 99), es.max())


@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("R", [1, 127, 128, 129, 300, 4096 + 77])
def test_mlp_tensor_core_ragged_rows_vs_fp32_kernel(precision, R):
    """Row counts around the 128-row tile and the 2-tile slot pairing, vs the on-device fp32 path."""
    rng = np.random.default_rng(R)
    w = om.init_weights(3, bias_scale=0.05)
    nerf = make_nerf(w, precision)
    xyz = dev(rng.uniform(-1, 1, (R, 3)).astype(F32))
    d = rng.normal(size=(R, 3)); d = dev((d / np.linalg.norm(d, axis=1, keepdims=True)).astype(F32))
    tol = TC_TOL[precision]
    for sub in (nerf.coarse_model, nerf.fine_model):
        rgb, sig = sub((xyz, d))
        rgb32, sig32 = sub((xyz, d), precision=_lib.FP32)
        e = (rgb - rgb32).abs()
        assert float(e.max()) <= tol["rgb_max"], float(e.max())
        es = (sig - sig32).abs() / torch.clamp(sig32.abs(), min=1.0)
        assert float(es.max()) <= tol["sig_rel_max"], float(es.max())
        assert torch.isfinite(rgb).all() and torch.isfinite(sig).all()


def test_forward_fp32_end_to_end_vs_oracle(golden):
    """NeRF.forward, perturbation off, fixed uniforms, fp32 MLP: every stage near-exact."""
    g = golden["oracle_forward_train"]
    for tag, gain in (("g1", 1.0), ("g300", 300.0)):
        nerf = make_nerf(om.init_weights(7, sigma_gain=gain), "fp32")
        pc, pf = nerf.forward(dev(g["rays_o"]), dev(g["rays_d"]), dev(g["near"]), dev(g["far"]), u_fine=dev(g["u_fine"]))
        assert set(pf) == {"acc_map", "weights", "pred_rgb", "pred_depth"}
        assert pf["weights"].shape == (64, 192) and pc["weights"].shape == (64, 64)
        # fp32-vs-fp64 noise floor of the reference arithmetic is 2.4e-4
        assert np.abs(host(pc["pred_rgb"]) - g[f"{tag}_c_pred_rgb"]).max() <= 3e-4
        assert np.abs(host(pf["pred_rgb"]) - g[f"{tag}_f_pred_rgb"]).max() <= 5e-4
        assert np.abs(host(pf["pred_depth"]) - g[f"{tag}_f_pred_depth"]).max() <= 1e-3
        assert np.abs(host(pf["acc_map"]) - g[f"{tag}_f_acc_map"]).max() <= 1e-3


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_render_tensor_core_vs_oracle(precision):
    """predict() of a 40x40 synthetic 360-degree view (cfg1-shaped, reduced) vs the fp32 oracle:
    per-pixel absolute tolerance (p99 and bounded maximum) on rgb, depth and acc, PSNR-vs-GT within 0.1 dB."""
    H = W = 40
    v = osc.synthetic_view(H, W, view=1)
    rng = np.random.default_rng(11)
    uf = rng.random((H * W, 128), dtype=F32)
    w = om.init_weights(7)
    pc, pf = om.forward(w, v["rays_o"], v["rays_d"], v["near"], v["far"], u_fine=uf, perturb=False, white_bg=True)
    nerf = make_nerf(w, precision)
    ds = nb.RayDataset.from_tensor_slices(((v["rays_o"], v["rays_d"], v["near"], v["far"]),)).batch(512)
    # fixed uniforms: go through render_rays (predict draws Philox uniforms like the reference draws tf.random)
    oc, of = nerf.render_rays(dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]), u_fine=dev(uf), need_weights=True)
    # Stated tolerance (oracle/tolerance.py): per-pixel absolute, p99 and a BOUNDED maximum over every pixel of rgb, depth
    # and acc, coarse and fine; the only pixels exempt are the counted (<= 0.05 %) rays whose last-sample alpha sits on
    # the other side of the reference's 0/1 jump (delta_last = 1e10, utils/ray_utils.py:459-468).
    from oracle.tolerance import check_render
    npd = lambda d: {k: host(x) for k, x in d.items()}
    check_render(precision, npd(oc), npd(of), pc, pf)
    gt = rng.random((H * W, 3), dtype=F32)
    clip = lambda a: np.clip(a * 255.0, 0.0, 255.0) / 255.0
    assert abs(rm.psnr_metric_numpy(gt, clip(host(of["pred_rgb"]))) - rm.psnr_metric_numpy(gt, clip(pf["pred_rgb"]))) <= 0.1
    # the public predict() surface: structure, shapes, dtypes as Keras returns them
    out = nerf.predict(x=ds)
    assert isinstance(out, tuple) and len(out) == 2
    for d, S in zip(out, (64, 192)):
        assert d["pred_rgb"].shape == (H * W, 3) and d["weights"].shape == (H * W, S)
        assert d["pred_depth"].shape == (H * W,) and d["acc_map"].shape == (H * W,)
        assert all(isinstance(a, np.ndarray) and a.dtype == np.float32 for a in d.values())


def test_train_step_fp32_vs_oracle(golden):
    g = golden["oracle_forward_train"]
    nerf = make_nerf(om.init_weights(7), "fp32")
    batch = ((g["rays_o"], g["rays_d"], g["near"], g["far"]), (g["rgb_gt"],))
    logs = nerf.train_step(batch, u_fine=dev(g["u_fine"]))
    names = om.all_variable_names()
    assert abs(float(nerf.last_loss.item()) - float(g["train_loss"])) <= 2e-5 * float(g["train_loss"]) + 1e-6
    assert abs(logs["psnr_metric"] - float(g["train_psnr_metric"])) <= 2e-3
    gn = np.array([float(torch.linalg.vector_norm(nerf.flat_grads[v._ofs:v._ofs + v._n].double())) for v in nerf.trainable_variables])
    ref = g["train_grad_norms"]
    # fp32 kernels vs fp32 autograd: both are ~5e-4 (global) from the fp64 gradient
    assert np.all(np.abs(gn - ref) <= 3e-2 * ref + 1e-7), np.max(np.abs(gn - ref) / (ref + 1e-12))
    var = {v.name: v for v in nerf.trainable_variables}
    take = lambda n: host(nerf.flat_grads[var[n]._ofs:var[n]._ofs + var[n]._n]).reshape(var[n].shape)
    for n, key in (("fine/dense_9/bias", "train_grad_fine_dense_9_bias"), ("coarse/rgb/kernel", "train_grad_coarse_rgb_kernel"),
                   ("fine/sigma/kernel", "train_grad_fine_sigma_kernel")):
        a, b = take(n), g[key]
        assert np.abs(a - b).max() <= 2e-3 * np.abs(b).max() + 1e-8, n
    # Adam: the very first step moves every weight with a non-zero gradient by ~lr
    assert np.allclose(var["fine/rgb/kernel"].numpy(), g["train_param_after_fine_rgb_kernel"], atol=2e-5)
    assert nerf.optimizer.iterations == 1
    ov = nerf.optimizer.variables()
    assert len(ov) == 1 + 48 + 48 and ov[1].shape == (63, 256)


def test_fit_evaluate_surface_and_loss_decreases():
    H = W = 16
    v = osc.synthetic_view(H, W, view=0)
    rng = np.random.default_rng(0)
    gt = np.tile(np.array([[0.2, 0.5, 0.8]], F32), (H * W, 1))
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": False}}), precision="fp32", seed=1)
    ds = nb.RayDataset.from_tensor_slices(((v["rays_o"], v["rays_d"], v["near"], v["far"]), (gt,))).shuffle(seed=0).repeat().batch(128, drop_remainder=True)
    val = nb.RayDataset.from_tensor_slices(((v["rays_o"], v["rays_d"], v["near"], v["far"]), (gt,))).batch(128)
    before = nerf.evaluate(val)
    hist = nerf.fit(x=ds, epochs=3, steps_per_epoch=8, validation_data=val, validation_freq=3)
    after = nerf.evaluate(val)
    assert after > before + 1.0, (before, after)
    assert len(hist.history["psnr_metric"]) == 3 and "val_psnr_metric" in hist.history and "loss" in hist.history
    assert nerf.optimizer.iterations == 24


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_train_step_tensor_core_vs_oracle(golden, precision):
    """Tensor-core training step (forward stash -> backward-data -> weight-gradient kernels -> fused Adam)
    vs the fp32 autograd oracle. Stated tolerance (SURVEY.md App. E3 form): loss within 5e-3 relative,
    global gradient cosine >= 0.999, every one of the 48 gradient tensors cosine >= 0.98."""
    g = golden["oracle_forward_train"]
    nerf = make_nerf(om.init_weights(7), precision, train_precision=precision)
    ref = make_nerf(om.init_weights(7), "fp32")
    dv = lambda k: dev(g[k])
    args = (dv("rays_o"), dv("rays_d"), dv("near"), dv("far"), dv("rgb_gt"))
    loss, _, _ = nerf._loss_and_grads(*args, u_fine=dv("u_fine"))
    loss32, _, _ = ref._loss_and_grads(*args, u_fine=dv("u_fine"))
    assert abs(float(loss.item()) - float(g["train_loss"])) <= 5e-3 * float(g["train_loss"])
    a, b = nerf.flat_grads.double(), ref.flat_grads.double()
    assert torch.isfinite(a).all()
    assert float(torch.nn.functional.cosine_similarity(a, b, dim=0)) >= 0.999
    for v in nerf.trainable_variables:
        x, y = a[v._ofs:v._ofs + v._n], b[v._ofs:v._ofs + v._n]
        assert float(torch.nn.functional.cosine_similarity(x, y, dim=0)) >= 0.98, v.name
    # gradient norms also agree with the fp32 autograd oracle's
    gn = np.array([float(torch.linalg.vector_norm(a[v._ofs:v._ofs + v._n])) for v in nerf.trainable_variables])
    assert np.all(np.abs(gn - g["train_grad_norms"]) <= 0.15 * g["train_grad_norms"] + 1e-7)
    # and the full step runs through the public surface
    logs = nerf.train_step(((g["rays_o"], g["rays_d"], g["near"], g["far"]), (g["rgb_gt"],)), u_fine=dv("u_fine"))
    assert nerf.optimizer.iterations == 1 and np.isfinite(logs["psnr_metric"])


def test_tensor_core_training_reduces_loss():
    H = W = 16
    v = osc.synthetic_view(H, W, view=0)
    gt = np.tile(np.array([[0.2, 0.5, 0.8]], F32), (H * W, 1))
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": False}}), precision="bf16", train_precision="bf16", seed=1)
    ds = nb.RayDataset.from_tensor_slices(((v["rays_o"], v["rays_d"], v["near"], v["far"]), (gt,))).shuffle(seed=0).repeat().batch(128, drop_remainder=True)
    val = nb.RayDataset.from_tensor_slices(((v["rays_o"], v["rays_d"], v["near"], v["far"]), (gt,))).batch(128)
    before = nerf.evaluate(val)
    nerf.fit(x=ds, epochs=2, steps_per_epoch=16)
    after = nerf.evaluate(val)
    assert after > before + 1.0, (before, after)


@pytest.mark.parametrize("n_dw", [14, 48, 100])
def test_overlapped_backward_equals_sequential_backward(golden, n_dw):
    """The phase-split backward (coarse weight-gradient phase on `n_dw` SMs and a second stream, next to the fine
    backward-data phase on the rest) computes the same gradients as the sequential one. Same kernels, same operands;
    only the K-split of the weight-gradient GEMM (hence the fp32 summation order) depends on the SM count."""
    g = golden["oracle_forward_train"]
    dv = lambda k: dev(g[k])
    args = (dv("rays_o"), dv("rays_d"), dv("near"), dv("far"), dv("rgb_gt"))
    grads = []
    for n in (0, n_dw):
        nerf = make_nerf(om.init_weights(7), "bf16", train_precision="bf16")
        nerf._dw_overlap_sms = n
        for _ in range(2):                                  # twice: the second pass reuses stashes, streams, workspaces
            loss, _, _ = nerf._loss_and_grads(*args, u_fine=dv("u_fine"))
        torch.cuda.synchronize()
        grads.append((float(loss.item()), nerf.flat_grads.double().clone()))
    (l0, a), (l1, b) = grads
    assert abs(l0 - l1) <= 1e-6 * abs(l0)          # (the loss is a sum of per-ray float atomics: equal to rounding)
    scale = float(a.abs().max())
    assert float((a - b).abs().max()) <= 2e-5 * scale
    assert float(torch.nn.functional.cosine_similarity(a, b, dim=0)) >= 1 - 1e-9


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("B,S", [(1, 2), (127, 3), (300, 64), (1000, 192)])
def test_split_last_sample_rows_match_fp32_sigma(precision, B, S):
    """The split-operand launch (NERFB200_OPT_PRECISE_LAST) rewrites sigma of row ray*S + S-1 of every ray and nothing
    else: those rows agree with the on-device fp32 path to fp32-grade accuracy (hi.Whi + lo.Whi + hi.Wlo), the other
    rows are bit-identical to a forward without it, ragged ray counts around the 128-row tile included."""
    rng = np.random.default_rng(B * 1000 + S)
    w = om.init_weights(3, bias_scale=0.05)
    ro = dev(rng.uniform(-0.3, 0.3, (B, 3)).astype(F32))
    d = rng.normal(size=(B, 3)); rd = dev((d / np.linalg.norm(d, axis=1, keepdims=True)).astype(F32))
    t = dev(np.sort(rng.uniform(0.4, 1.3, (B, S)).astype(F32), axis=1))
    on, off = make_nerf(w, precision, precise_last=True), make_nerf(w, precision, precise_last=False)
    for which in (0, 1):
        rgb1, s1 = on._mlp(which, ro, rd, t)
        rgb0, s0 = off._mlp(which, ro, rd, t)
        _, s32 = on._mlp(which, ro, rd, t, precision=_lib.FP32)
        s1, s0, s32 = (x.reshape(B, S) for x in (s1, s0, s32))
        assert torch.equal(rgb1, rgb0) and torch.equal(s1[:, :-1], s0[:, :-1])
        scale = torch.clamp(s32[:, -1].abs(), min=1.0)
        e_split = float(((s1[:, -1] - s32[:, -1]).abs() / scale).max())
        e_plain = float(((s0[:, -1] - s32[:, -1]).abs() / scale).max())
        lim = 5e-4 if precision == "bf16" else 1e-4
        assert e_split <= lim, (e_split, e_plain)
        if B >= 100:
            assert e_split < 0.1 * e_plain, (e_split, e_plain)
        # the sign of the ReLU'd sigma agrees with fp32 wherever fp32 is not within the split launch's own error of 0
        clear = s32[:, -1] > 3 * lim
        assert bool(((s1[:, -1] > 0) == (s32[:, -1] > 0))[clear].all())


def test_tf32_is_a_render_precision_only():
    nerf = make_nerf(om.init_weights(3), "tf32")
    assert nerf.train_precision == _lib.BF16
    lib = _lib.load()
    z = torch.zeros(8, device="cuda")
    assert lib.nerfb200_mlp_backward(nerf._ctx, 0, 1, 1, _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), _lib.ptr(z),
                                     _lib.ptr(z), _lib.TF32, None, None, None) == 10002
    assert lib.nerfb200_mlp_stash_bytes(1000, _lib.TF32) == 0


def test_context_is_bound_to_its_device():
    """Constant tables are uploaded per device and launches are refused from another current device (the Python
    surface enters the model's device itself)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = om.init_weights(3)
    rng = np.random.default_rng(0)
    xyz = rng.uniform(-1, 1, (300, 3)).astype(F32)
    dd = rng.normal(size=(300, 3)); dd = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(F32)
    outs = []
    for di in (0, 1):
        p = nb.make_params({"system": {"white_bg": True}}, perturb=False)
        nerf = nb.setup_model(p, precision="bf16", device=f"cuda:{di}")
        nerf.set_weights_from_dict(w)
        x, dv_ = torch.from_numpy(xyz).to(f"cuda:{di}"), torch.from_numpy(dd).to(f"cuda:{di}")
        rgb, sig = nerf.coarse_model((x, dv_))          # current device stays 0: the model enters its own device
        outs.append((rgb.cpu(), sig.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_graphed_train_step_equals_eager(precision):
    """train_step captured as ONE CUDA graph (device-resident step state: sampling step and Adam iteration advance
    inside the graph) leaves bit-identical parameters, optimizer state, loss and metric as the eager launches, with the
    in-kernel sampling noise ON (so a replay that reused the captured step's noise or learning rate would differ)."""
    H = W = 32
    v = osc.synthetic_view(H, W, view=0)
    rng = np.random.default_rng(0)
    gt = rng.random((H * W, 3), dtype=F32)
    batches = [tuple(dev(v[k][s0:s0 + 256]) for k in ("rays_o", "rays_d", "near", "far")) + (dev(gt[s0:s0 + 256]),)
               for s0 in (0, 256, 512)]
    finals = []
    for graph in (False, True):
        nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=True), precision="bf16",
                              train_precision=precision, seed=2, rng_seed=7, cuda_graph=graph)
        logs = None
        for i in range(7):
            b = batches[i % 3]
            logs = nerf.train_step(((b[0], b[1], b[2], b[3]), (b[4],)))
        torch.cuda.synchronize()
        if graph:
            assert any("graph" in st for st in nerf._graphs.values()), "the step was never captured"
        finals.append((nerf.flat_params.clone(), nerf.optimizer.m.clone(), nerf.optimizer.v.clone(), float(nerf.last_loss.item()),
                       float(logs["psnr_metric"]), nerf.optimizer.iterations))
    (p0, m0, v0, l0, q0, it0), (p1, m1, v1, l1, q1, it1) = finals
    assert it0 == it1 == 7
    if precision == "bf16":        # deterministic kernels: the replayed step is the eager step, bit for bit
        assert torch.equal(p0, p1) and torch.equal(m0, m1) and torch.equal(v0, v1)
        # (the loss value and the metric state are sums of float atomics over the blocks of the loss kernel: equal to
        # rounding, not bit for bit -- the gradient does not depend on them)
        assert abs(l0 - l1) <= 1e-6 * abs(l0) and abs(q0 - q1) <= 1e-5
    else:                          # the fp32 check path accumulates weight gradients with atomics (order varies run to run)
        # (and Adam's first steps move a weight by ~lr * sign(g): a gradient at the noise level may move the other way)
        d = (p0 - p1).abs()
        assert float(d.mean()) <= 2e-6 and float(d.max()) <= 7 * 2 * 5e-4 and abs(l0 - l1) <= 1e-4 * abs(l0) and abs(q0 - q1) <= 1e-2
    # and the operand images were repacked inside the graph: a render right after uses the new weights
    assert nerf._dirty is False


def test_split_launch_in_training_forwards(golden):
    """NERFB200_OPT_PRECISE_LAST = 2 (`precise_last="train"`): the training forward takes the split launch too; the
    patched sigma also lands in the activation stash (the ReLU gate of the sigma head in backward-data reads it), so
    loss and gradients stay consistent with each other and with the oracle."""
    g = golden["oracle_forward_train"]
    dv = lambda k: dev(g[k])
    args = (dv("rays_o"), dv("rays_d"), dv("near"), dv("far"), dv("rgb_gt"))
    out = {}
    for mode in (True, "train"):
        nerf = make_nerf(om.init_weights(7), "bf16", train_precision="bf16", precise_last=mode)
        loss, pp_c, pp_f = nerf._loss_and_grads(*args, u_fine=dv("u_fine"))
        torch.cuda.synchronize()
        out[mode] = (float(loss.item()), nerf.flat_grads.double().clone(), pp_f["pred_rgb"].clone())
        assert torch.isfinite(nerf.flat_grads).all()
        assert abs(out[mode][0] - float(g["train_loss"])) <= 5e-3 * float(g["train_loss"])
    # the rendered pixels of the training forward with the split launch equal the render forward's (same kernels + launch)
    ref = make_nerf(om.init_weights(7), "bf16")
    _, rf = ref.forward(dv("rays_o"), dv("rays_d"), dv("near"), dv("far"), u_fine=dv("u_fine"))
    assert float((out["train"][2] - rf["pred_rgb"]).abs().max()) <= 1e-6
    cos = float(torch.nn.functional.cosine_similarity(out[True][1], out["train"][1], dim=0))
    assert cos >= 0.99, cos
