"""GPU parity tests of the HBM-bound kernels (rays, samplers, integrator) through the C ABI,
against the CPU oracle and the committed reference fixtures. Integer/index work is bit-exact;
floating point is within the tolerance written next to each assert."""
import numpy as np
import pytest
import torch

import nerf_tf2_b200 as nb
from nerf_tf2_b200 import ray_utils as ru
from oracle import ray_march as rm, scene as osc

pytestmark = pytest.mark.gpu
F32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def test_get_rays_matches_reference_fixture(golden):
    g = golden["ref_rays"]
    for tag in ("a", "b"):
        H, W = int(g[f"{tag}_H"]), int(g[f"{tag}_W"])
        ro, rd = ru.get_rays(H, W, g[f"{tag}_K"], g[f"{tag}_c2w"])
        assert np.array_equal(host(ro), g[f"{tag}_rays_o"])
        # fp64 maths rounded to fp32: at most 1 ulp (numpy's BLAS may fuse the 3-term dot differently)
        assert np.abs(host(rd) - g[f"{tag}_rays_d"]).max() <= 6e-8
        ro2, rd2 = ru.get_rays_tf(H, W, g[f"{tag}_K"], g[f"{tag}_c2w"])
        assert np.abs(host(rd2) - g[f"{tag}_rays_d_tf"]).max() <= 1.2e-7
        # ray sharding: a sub-range equals the slice of the full image
        ro3, rd3 = ru.get_rays(H, W, g[f"{tag}_K"], g[f"{tag}_c2w"], ray0=W + 1, n_rays=2 * W)
        assert torch.equal(rd3, rd[W + 1:3 * W + 1])
        ids = torch.tensor([0, 5, H * W - 1, 7], dtype=torch.int32).cuda()
        ro4, rd4 = ru.get_rays_at(H, W, g[f"{tag}_K"], g[f"{tag}_c2w"], ids)
        assert torch.equal(rd4, rd2[ids.long()])


@pytest.mark.parametrize("tag,lin_inv", [("inv", True), ("lin", False)])
def test_coarse_sampler_bit_exact(golden, tag, lin_inv):
    g = golden["ref_sampling_composite"]
    t, e = ru.sample_coarse(64, lin_inv, True, dev(g["near"]), dev(g["far"]), dev(g[f"{tag}_u_coarse"]))
    assert np.array_equal(host(e), g[f"{tag}_bin_edges"])
    assert np.array_equal(host(t), g[f"{tag}_t_coarse"])
    xyz, dirs = ru.make_inputs(dev(g["rays_o"]), dev(g["rays_d"]), t)
    assert np.array_equal(host(xyz), g[f"{tag}_xyz_coarse"])
    # perturb off (reference crashes; oracle = mid-points)
    t0, e0 = ru.sample_coarse(64, lin_inv, False, dev(g["near"]), dev(g["far"]))
    o = rm.create_input_batch_coarse_model(64, lin_inv, False, g["rays_o"], g["rays_d"], g["near"], g["far"])
    assert np.array_equal(host(t0), o["t_vals"])


def test_coarse_sampler_philox_uniform_and_shard_invariant():
    B = 4096
    near = torch.full((B,), 0.425).cuda(); far = torch.full((B,), 1.275).cuda()
    t, e = ru.sample_coarse(64, True, True, near, far, None, seed=5, ray0=0)
    u = (t - e[:, :-1]) / (e[:, 1:] - e[:, :-1])
    assert 0 <= float(u.min()) and float(u.max()) < 1.0 + 1e-5 and abs(float(u.mean()) - 0.5) < 5e-3
    t2, _ = ru.sample_coarse(64, True, True, near[1000:2000].contiguous(), far[1000:2000].contiguous(), None, seed=5, ray0=1000)
    assert torch.equal(t2, t[1000:2000])
    t3, _ = ru.sample_coarse(64, True, True, near, far, None, seed=6, ray0=0)
    assert not torch.equal(t3, t)


@pytest.mark.parametrize("tag", ["inv", "lin"])
def test_integrator_matches_reference_fixture(golden, tag):
    g = golden["ref_sampling_composite"]
    for wb in (0, 1):
        pp = ru.post_process_model_output(dev(g[f"{tag}_rgb"]), dev(g[f"{tag}_sigma"]), dev(g[f"{tag}_t_coarse"]), bool(wb))
        # fp32: expf is within 2 ulp and the transmittance product is a tree scan, not sequential
        assert np.allclose(host(pp["weights"]), g[f"{tag}_wb{wb}_weights"], rtol=2e-5, atol=1e-7)
        assert np.allclose(host(pp["pred_rgb"]), g[f"{tag}_wb{wb}_pred_rgb"], rtol=1e-5, atol=2e-6)
        assert np.allclose(host(pp["pred_depth"]), g[f"{tag}_wb{wb}_pred_depth"], rtol=1e-5, atol=2e-6)
        assert np.allclose(host(pp["acc_map"]), g[f"{tag}_wb{wb}_acc_map"], rtol=1e-5, atol=2e-6)
    ppf = ru.post_process_model_output(dev(g[f"{tag}_rgb_f"]), dev(g[f"{tag}_sigma_f"]), dev(g[f"{tag}_t_fine_sorted"]), True)
    for k in ("weights", "pred_rgb", "pred_depth", "acc_map"):
        assert np.allclose(host(ppf[k]), g[f"{tag}_fine_wb1_{k}"], rtol=2e-5, atol=2e-6), k
    w = ru.compute_weights(dev(g[f"{tag}_sigma"]), dev(g[f"{tag}_t_coarse"]))
    assert np.allclose(host(w), g[f"{tag}_wb0_weights"], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("S", [2, 32, 33, 64, 96, 100, 128, 192, 256, 384, 512, 1000, 1024])
def test_integrator_ragged_sample_counts(S):
    rng = np.random.default_rng(S)
    B = 37
    t = np.sort(rng.random((B, S), dtype=F32) * F32(0.8) + F32(0.4), axis=1)
    sig = (rng.random((B * S, 1), dtype=F32) * 20 * (rng.random((B * S, 1)) > 0.5)).astype(F32)
    rgb = rng.random((B * S, 3), dtype=F32)
    o = rm.post_process_model_output(rgb, sig, t, True)
    pp = ru.post_process_model_output(dev(rgb), dev(sig), dev(t), True)
    for k in ("weights", "pred_rgb", "pred_depth", "acc_map"):
        assert np.allclose(host(pp[k]), o[k], rtol=5e-5, atol=3e-6), k


def test_integrator_empty_batch_and_no_weights():
    z = torch.zeros((0, 64)).cuda()
    pp = ru.post_process_model_output(torch.zeros((0, 3)).cuda(), torch.zeros((0,)).cuda(), z, True)
    assert pp["pred_rgb"].shape == (0, 3)
    rng = np.random.default_rng(0)
    t = np.sort(rng.random((8, 64), dtype=F32), axis=1); sig = rng.random((512, 1), dtype=F32); rgb = rng.random((512, 3), dtype=F32)
    a = ru.post_process_model_output(dev(rgb), dev(sig), dev(t), False, need_weights=False)
    b = ru.post_process_model_output(dev(rgb), dev(sig), dev(t), False)
    assert "weights" not in a and torch.equal(a["pred_rgb"], b["pred_rgb"])


def test_integrator_backward_matches_autograd():
    from oracle import model as om
    rng = np.random.default_rng(4)
    for S, wb in ((64, True), (192, False), (96, True)):
        B = 29
        t = np.sort(rng.random((B, S), dtype=F32) * F32(0.8) + F32(0.4), axis=1)
        sig = (rng.random((B * S,), dtype=F32) * 15 * (rng.random((B * S,)) > 0.4)).astype(F32)
        rgb = rng.random((B * S, 3), dtype=F32)
        dC = rng.normal(size=(B, 3)).astype(F32)
        st = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
        rt = torch.tensor(rgb, dtype=torch.float64, requires_grad=True)
        pp = om.composite_torch(rt, st, torch.tensor(t, dtype=torch.float64), wb)
        (pp["pred_rgb"] * torch.tensor(dC, dtype=torch.float64)).sum().backward()
        ds, dr = ru.composite_backward(dev(rgb), dev(sig), dev(t), wb, dev(dC))
        # fp32 kernel vs fp64 autograd; the last sample (delta = 1e10) is excluded where sigma == 0:
        # its true derivative is +-1e10-scaled and meaningless for parity
        g_ref = st.grad.numpy().reshape(B, S); g_k = host(ds).reshape(B, S)
        mask = np.ones((B, S), bool); mask[:, -1] = False
        scale = np.abs(g_ref[mask]).max()
        assert np.abs(g_k[mask] - g_ref[mask]).max() <= 2e-4 * scale
        assert np.allclose(host(dr), rt.grad.numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("S", [64, 192, 384, 100, 33, 2])
@pytest.mark.parametrize("white_bg", [True, False])
def test_training_integrator_equals_forward_loss_backward_launches(S, white_bg):
    """nerfb200_composite_train (integrator + MeanSquaredError term + integrator backward as ONE launch, NeRF.train_step,
    core/model.py:148-170) against the three launches it replaces: forward outputs BIT-IDENTICAL to
    nerfb200_composite_fwd (same lane mapping and operations, so a training forward resamples exactly like a render
    forward), gradients equal to nerfb200_composite_bwd fed with nerfb200_mse_loss_grad's d_pred up to fp32 summation
    order, loss and metric state equal to rounding; odd ray counts (the two-rays-per-warp form of S = 64 with a last
    half-empty warp), rows that are not 16-byte aligned, accumulation into a non-zero loss."""
    from nerf_tf2_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(S)
    for B, Bg in ((37, 37), (256, 1024), (1, 1)):
        t = np.sort(rng.random((B, S), dtype=F32) * F32(0.8) + F32(0.4), axis=1)
        sig = (rng.random((B * S,), dtype=F32) * 15 * (rng.random((B * S,)) > 0.4)).astype(F32)
        rgb = rng.random((B * S, 3), dtype=F32)
        gt = rng.random((B, 3), dtype=F32)
        dt, dsg, drgb, dgt = dev(t), dev(sig), dev(rgb), dev(gt)
        # the three launches
        pp = ru.post_process_model_output(drgb, dsg, dt, white_bg)
        d_pred = torch.empty((B, 3), device="cuda"); loss0 = torch.full((1,), 0.25, device="cuda"); met0 = torch.zeros(2, device="cuda")
        _lib.check(lib.nerfb200_mse_loss_grad(B, Bg, _lib.ptr(pp["pred_rgb"]), _lib.ptr(dgt), _lib.ptr(d_pred), _lib.ptr(loss0),
                                              _lib.ptr(met0), _lib.stream_ptr()), "mse_loss_grad")
        ds0, dr0 = ru.composite_backward(drgb, dsg, dt, white_bg, d_pred)
        # the one launch
        loss1 = torch.full((1,), 0.25, device="cuda"); met1 = torch.zeros(2, device="cuda")
        pp1, ds1, dr1 = ru.composite_train(drgb, dsg, dt, white_bg, dgt, Bg, loss1, metric_state=met1)
        for k in ("pred_rgb", "pred_depth", "acc_map", "weights"):
            assert torch.equal(pp[k], pp1[k]), (S, B, k)
        assert abs(float(loss1) - float(loss0)) <= 2e-6 * abs(float(loss0))
        assert abs(float(met1[0]) - float(met0[0])) <= 2e-6 * max(1.0, float(met0[0])) and float(met1[1]) == float(met0[1]) == B
        m = torch.ones((B, S), dtype=torch.bool, device="cuda"); m[:, -1] = False       # delta_last = 1e10 amplifies rounding
        scale = float(ds0.reshape(B, S)[m].abs().max())
        assert float((ds1.reshape(B, S) - ds0.reshape(B, S))[m].abs().max()) <= 2e-5 * max(scale, 1e-30), (S, B)
        last0, last1 = ds0.reshape(B, S)[:, -1], ds1.reshape(B, S)[:, -1]
        assert torch.allclose(last1, last0, rtol=1e-3, atol=1e-3 * float(last0.abs().max()) + 1e-30)
        assert torch.allclose(dr1, dr0, rtol=1e-6, atol=1e-9)
        # without the optional outputs; the same forward results
        loss2 = torch.zeros(1, device="cuda")
        pp2, ds2, dr2 = ru.composite_train(drgb, dsg, dt, white_bg, dgt, Bg, loss2, need_weights=False)
        assert "weights" not in pp2 and torch.equal(pp2["pred_rgb"], pp["pred_rgb"]) and torch.equal(ds2, ds1) and torch.equal(dr2, dr1)
    # rows that are not 16-byte aligned take the scalar staging path: identical results
    B = 9
    t = np.sort(rng.random((B, S), dtype=F32), axis=1); sig = rng.random((B * S,), dtype=F32); rgb = rng.random((B * S, 3), dtype=F32)
    gt = dev(rng.random((B, 3), dtype=F32))
    pad = lambda a: torch.cat([torch.zeros(1, device="cuda"), dev(a).reshape(-1)])[1:].reshape(a.shape)      # 4-byte offset view
    la, lb = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    pa, dsa, dra = ru.composite_train(dev(rgb), dev(sig), dev(t), white_bg, gt, B, la)
    tt, ss, rr = pad(t), pad(sig), pad(rgb)
    assert tt.data_ptr() % 16 != 0
    out = {k: torch.empty_like(v) for k, v in pa.items()}
    dsb, drb = torch.empty_like(dsa), torch.empty_like(dra)
    _lib.check(lib.nerfb200_composite_train(B, S, ss.data_ptr(), rr.data_ptr(), tt.data_ptr(), int(white_bg), _lib.ptr(gt), B,
                                            _lib.ptr(out["weights"]), _lib.ptr(out["pred_rgb"]), _lib.ptr(out["pred_depth"]),
                                            _lib.ptr(out["acc_map"]), _lib.ptr(dsb), _lib.ptr(drb), _lib.ptr(lb), None,
                                            _lib.stream_ptr()), "composite_train")
    for k in pa:
        assert torch.equal(pa[k], out[k]), k
    assert torch.equal(dsa, dsb) and torch.equal(dra, drb) and abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))   # (float atomics)
    # argument validation
    assert lib.nerfb200_composite_train(4, 1, None, None, None, 0, None, 4, None, None, None, None, None, None, None, None, None) == 10001
    assert lib.nerfb200_composite_train(4, 64, 1, 1, 1, 0, 1, 2, None, 1, 1, 1, 1, 1, 1, None, None) == 10001     # B_global < B


@pytest.mark.parametrize("tag", ["inv", "lin"])
def test_fine_sampler_vs_reference_fixture(golden, tag):
    g = golden["ref_sampling_composite"]
    w = g[f"{tag}_wb0_weights"]
    ts, dbg = ru.sample_fine(128, dev(w), dev(g[f"{tag}_bin_edges"]), dev(g[f"{tag}_t_coarse"]),
                             dev(g[f"{tag}_u_fine"]), debug=True)
    ts, cdf, idx, tfine = host(ts), host(dbg["cdf"]), host(dbg["piece_idxs"]), host(dbg["t_vals_fine"])
    assert idx.dtype == np.int32
    # (1) searchsorted indices are BIT-EXACT on the same fp32 CDF (north-star contract)
    assert np.array_equal(idx, rm.searchsorted_right(cdf[:, 1:-1], g[f"{tag}_u_fine"]))
    # (2) the CDF itself agrees with the reference's sequential-sum CDF to accumulated fp32 rounding
    # (64 additions near 1.0, ulp 1.2e-7; the kernel sums lane-serial + shuffle-scan)
    assert np.abs(cdf - g[f"{tag}_cdf"]).max() <= 2e-6
    # (3) indices vs the reference's own CDF: identical except where u sits within rounding of an edge
    assert (idx != g[f"{tag}_piece_idxs"]).mean() <= 2e-3
    # (4) inversion given the kernel's own cdf/idx reproduces the reference formula exactly
    left = g[f"{tag}_bin_edges"][:, :-1]
    widths = g[f"{tag}_bin_edges"][:, 1:] - left
    pdf, _ = rm.fine_cdf(w, widths)
    # (5) output is the ascending sort of concat(t_coarse, t_fine): exact as a multiset
    assert np.array_equal(ts, np.sort(np.concatenate([g[f"{tag}_t_coarse"], tfine], axis=1), axis=1))
    # (6) and within fp32 rounding of the reference's sorted samples (t in [0.4, 1.3])
    # (t = (u - cdf)/pdf + left amplifies the ~1e-6 CDF rounding by 1/pdf where the pdf is tiny)
    assert np.abs(ts - g[f"{tag}_t_fine_sorted"]).max() <= 2e-4
    close = np.abs(ts - g[f"{tag}_t_fine_sorted"]) <= 2e-6
    assert close.mean() >= 0.99


def test_fine_sampler_edge_cases_and_shapes():
    rng = np.random.default_rng(1)
    for Nc, Nf in ((64, 128), (128, 256), (32, 32), (96, 64), (256, 512)):
        B = 19
        near = np.full((B, 1), 0.425, F32); far = np.full((B, 1), 1.275, F32)
        o = rm.create_input_batch_coarse_model(Nc, True, True, np.zeros((B, 3), F32), np.ones((B, 3), F32), near, far,
                                               rng.random((B, Nc), dtype=F32))
        w = rng.random((B, Nc), dtype=F32) ** 8
        w[0] = 0.0                       # all-zero weights
        w[1] = 0.0; w[1, Nc // 3] = 1.0  # one-hot
        u = rng.random((B, Nf), dtype=F32)
        u[2, 0] = 0.0; u[2, 1] = np.nextafter(F32(1), F32(0))
        ts, dbg = ru.sample_fine(Nf, dev(w), dev(o["bin_data"]["bin_edges"]), dev(o["t_vals"]), dev(u), debug=True)
        ts, cdf, idx, tf = host(ts), host(dbg["cdf"]), host(dbg["piece_idxs"]), host(dbg["t_vals_fine"])
        assert ts.shape == (B, Nc + Nf)
        assert np.array_equal(idx, rm.searchsorted_right(cdf[:, 1:-1], u))
        assert idx.min() >= 0 and idx.max() <= Nc - 1
        assert np.array_equal(ts, np.sort(np.concatenate([o["t_vals"], tf], axis=1), axis=1))
        ref = rm.create_input_batch_fine_model(np.zeros((B, 3), F32), np.ones((B, 3), F32), w, o["bin_data"], o["t_vals"], u,
                                               return_debug=True)
        assert np.abs(cdf - ref["cdf"]).max() <= 4e-6
        assert (np.abs(ts - ref["t_vals"]) <= 2e-6).mean() >= 0.995


def test_fine_sampler_unsorted_coarse_falls_back_to_full_sort():
    rng = np.random.default_rng(2)
    B, Nc, Nf = 9, 64, 128
    edges = rm.tf_linspace(np.full((B, 1), 0.4, F32), np.full((B, 1), 1.2, F32), Nc + 1)
    tc = rng.permuted(F32(0.5) * (edges[:, :-1] + edges[:, 1:]), axis=1)     # NOT ascending
    w = rng.random((B, Nc), dtype=F32); u = rng.random((B, Nf), dtype=F32)
    ts, dbg = ru.sample_fine(Nf, dev(w), dev(edges), dev(tc), dev(u), debug=True)
    assert np.array_equal(host(ts), np.sort(np.concatenate([tc, host(dbg["t_vals_fine"])], axis=1), axis=1))


def test_fine_sampler_philox_shard_invariant():
    rng = np.random.default_rng(3)
    B = 512
    edges = rm.tf_linspace(np.full((B, 1), 0.4, F32), np.full((B, 1), 1.2, F32), 65)
    tc = F32(0.5) * (edges[:, :-1] + edges[:, 1:])
    w = rng.random((B, 64), dtype=F32)
    a = ru.sample_fine(128, dev(w), dev(edges), dev(tc), None, seed=9, ray0=0)
    b = ru.sample_fine(128, dev(w[100:300]), dev(edges[100:300]), dev(tc[100:300]), None, seed=9, ray0=100)
    assert torch.equal(a[100:300], b)
    assert float(a.min()) >= 0.4 - 1e-6 and float(a.max()) <= 1.2 + 1e-5


def test_positional_encoding_and_depth_map(golden):
    g = golden["oracle_mlp"]
    enc = ru.positional_encode(dev(g["xyz"]), 10)
    # same fp32 argument x*fl32(2^l*pi); sinf/cosf are <= 2 ulp from libm's
    assert np.abs(host(enc) - g["enc_xyz_L10"]).max() <= 5e-7
    assert np.abs(host(ru.positional_encode(dev(g["dirs"]), 4)) - g["enc_dir_L4"]).max() <= 5e-7
    v = osc.synthetic_view(9, 11, view=2)
    depth = np.random.default_rng(0).random(99, dtype=F32) + F32(0.3)
    for mt in ("type_1", "type_2"):
        ref = rm.create_depth_map(depth, 9, 11, 0.2125, mt, v["K"], v["c2w"])
        out = ru.create_depth_map(dev(depth), 9, 11, 0.2125, mt, v["K"], v["c2w"])
        assert np.allclose(host(out), ref, rtol=2e-6, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# The placement fast path of the fine sampler (N_coarse/N_fine = 64/128 and 128/256): whatever the
# inputs, t_sorted must be the exact ascending sort of concat(t_coarse, t_fine) and the indices the
# exact upper bounds in the kernel's own CDF. Cases chosen to hit the self-check's fallback too.
def _fast_sampler_case(Nc, Nf, w, u, t_c=None, edges=None, seed=0):
    B = w.shape[0]
    rng = np.random.default_rng(seed)
    if edges is None:
        near = np.full((B, 1), 0.425, F32); far = np.full((B, 1), 1.275, F32)
        o = rm.create_input_batch_coarse_model(Nc, True, True, np.zeros((B, 3), F32), np.ones((B, 3), F32), near, far,
                                               rng.random((B, Nc), dtype=F32))
        edges = o["bin_data"]["bin_edges"]
        if t_c is None:
            t_c = o["t_vals"]
    ts, dbg = ru.sample_fine(Nf, dev(w), dev(edges), dev(t_c), dev(u), debug=True)
    ts, cdf, idx, tf = host(ts), host(dbg["cdf"]), host(dbg["piece_idxs"]), host(dbg["t_vals_fine"])
    assert np.array_equal(idx, rm.searchsorted_right(cdf[:, 1:-1], u))
    assert np.array_equal(ts, np.sort(np.concatenate([t_c, tf], axis=1), axis=1))
    # without the debug outputs the kernel takes the same path and returns the same samples
    assert np.array_equal(host(ru.sample_fine(Nf, dev(w), dev(edges), dev(t_c), dev(u))), ts)
    return ts, tf, idx


@pytest.mark.parametrize("Nc,Nf", [(64, 128), (128, 256)])
def test_fine_sampler_fast_path_bulk_exact(Nc, Nf):
    rng = np.random.default_rng(Nc)
    B = 4099                                            # not a multiple of the warps per CTA
    sharp = rng.choice([1, 4, 16, 64], size=(B, 1)).astype(F32)
    w = (rng.random((B, Nc), dtype=F32) ** sharp).astype(F32)      # flat ... nearly one-hot pdfs
    w[rng.random((B, Nc)) < 0.3] = 0.0                             # empty space (sigma == 0 for most samples)
    u = rng.random((B, Nf), dtype=F32)
    _fast_sampler_case(Nc, Nf, w, u, seed=1)


@pytest.mark.parametrize("Nc,Nf", [(64, 128), (128, 256)])
def test_fine_sampler_fast_path_adversarial(Nc, Nf):
    rng = np.random.default_rng(7)
    B = 24
    w = rng.random((B, Nc), dtype=F32)
    u = rng.random((B, Nf), dtype=F32)
    one_m = np.nextafter(F32(1), F32(0))
    u[0] = 0.0                                # every sample in one u-bucket, all t equal (ties by arrival slot)
    u[1] = 0.5
    u[2] = one_m                              # beyond cdf_last: last bin, t may pass `far` by ulps
    u[3] = np.sort(u[3]); u[4] = np.sort(u[4])[::-1]
    u[5, ::2] = u[5, 1::2]                    # duplicate pairs
    u[6] = (np.arange(Nf, dtype=F32) / F32(Nf))            # exactly on the bucket boundaries
    u[7] = np.repeat(rng.random(Nf // 8, dtype=F32), 8)    # 8-fold duplicates
    w[8] = 0.0                                # uniform pdf
    w[9] = 0.0; w[9, 5] = 1.0                 # one-hot: almost every sample in one bin
    w[10] = 0.0; w[10, Nc - 1] = 1e6          # pdf < 1e-8 elsewhere: masked inversion returns the left edge
    w[11] = 0.0; w[11, 0] = 1e6; u[11, :Nf // 2] = one_m   # ... with half the samples landing in masked bins
    w[12] = 1e-12
    _fast_sampler_case(Nc, Nf, w, u, seed=2)
    # coarse samples sitting exactly on the bin edges: ties between a coarse and a fine sample
    near = np.full((B, 1), 0.4, F32); far = np.full((B, 1), 1.2, F32)
    edges = rm.tf_linspace(near, far, Nc + 1)
    for t_c in (edges[:, :-1].copy(), edges[:, 1:].copy()):
        uu = rng.random((B, Nf), dtype=F32)
        uu[:, :8] = 0.0
        _fast_sampler_case(Nc, Nf, w, uu, t_c=t_c, edges=edges)
    # duplicate coarse samples (still ascending) and a descending row (generic path) in the same launch
    t_c = np.repeat(edges[:, :-1:2], 2, axis=1).copy()
    t_c[3] = t_c[3, ::-1]
    _fast_sampler_case(Nc, Nf, w, rng.random((B, Nf), dtype=F32), t_c=t_c, edges=edges)


def test_fine_sampler_unaligned_rows_take_the_generic_kernel():
    rng = np.random.default_rng(11)
    B, Nc, Nf = 33, 64, 128
    edges = rm.tf_linspace(np.full((B, 1), 0.4, F32), np.full((B, 1), 1.2, F32), Nc + 1)
    tc = F32(0.5) * (edges[:, :-1] + edges[:, 1:])
    w = rng.random((B, Nc), dtype=F32); u = rng.random((B, Nf), dtype=F32)
    a = ru.sample_fine(Nf, dev(w), dev(edges), dev(tc), dev(u))
    # same data at a 4-byte-offset address: not vector-aligned
    wbuf = torch.empty(B * Nc + 1, dtype=torch.float32, device="cuda")
    wbuf[1:] = dev(w).reshape(-1)
    b = ru.sample_fine(Nf, wbuf[1:].view(B, Nc), dev(edges), dev(tc), dev(u))
    assert torch.equal(a, b)


def test_integrator_unaligned_and_short_batches():
    rng = np.random.default_rng(12)
    for S in (64, 192):
        for B in (1, 7, 8, 9, 1025):
            t = np.sort(rng.random((B, S), dtype=F32) * F32(0.8) + F32(0.4), axis=1)
            sig = (rng.random((B * S,), dtype=F32) * 20 * (rng.random((B * S,)) > 0.5)).astype(F32)
            rgb = rng.random((B * S, 3), dtype=F32)
            a = ru.post_process_model_output(dev(rgb), dev(sig), dev(t), True)
            if B == 1025:
                o = rm.post_process_model_output(rgb, sig[:, None], t, True)
                for k in ("weights", "pred_rgb", "pred_depth", "acc_map"):
                    assert np.allclose(host(a[k]), o[k], rtol=5e-5, atol=3e-6), k
            # the same rows at addresses that are not 16-byte aligned: scalar staging, identical results
            sbuf = torch.empty(B * S + 1, dtype=torch.float32, device="cuda"); sbuf[1:] = dev(sig)
            rbuf = torch.empty(B * S * 3 + 3, dtype=torch.float32, device="cuda"); rbuf[3:] = dev(rgb).reshape(-1)
            b = ru.post_process_model_output(rbuf[3:].view(B * S, 3), sbuf[1:], dev(t), True)
            for k in ("weights", "pred_rgb", "pred_depth", "acc_map"):
                assert torch.equal(a[k], b[k]), (S, B, k)


@pytest.mark.parametrize("Nc,Nf,B", [(64, 128, 8192), (128, 256, 4096)])
def test_fine_sampler_in_kernel_uniforms_are_order_statistics(Nc, Nf, B):
    """Without explicit uniforms the fast path draws the ORDER STATISTICS of Nf i.i.d. U[0,1) directly (normalised
    partial sums of Nf+1 exponentials) instead of drawing Nf uniforms and sorting what they produce as the reference
    does (ray_utils.py:355,385): same joint distribution. Checked on a uniform pdf over linear bins, where
    t_fine = near + u * (far - near) up to rounding, against the Beta(k, Nf+1-k) moments of the k-th order statistic."""
    near, far = F32(0.4), F32(1.2)
    edges = rm.tf_linspace(np.full((B, 1), near, F32), np.full((B, 1), far, F32), Nc + 1)
    tc = F32(0.5) * (edges[:, :-1] + edges[:, 1:])
    w = np.ones((B, Nc), F32)
    ts, dbg = ru.sample_fine(Nf, dev(w), dev(edges), dev(tc), None, seed=21, debug=True)
    tf, idx = host(dbg["t_vals_fine"]).astype(np.float64), host(dbg["piece_idxs"])
    u = (tf - near) / (far - near)
    assert np.all(np.diff(u, axis=1) >= 0) and np.all(np.diff(idx, axis=1) >= 0)      # produced in order
    assert u.min() >= -1e-6 and u.max() < 1 + 1e-5
    assert np.array_equal(host(ts), np.sort(np.concatenate([tc, host(dbg["t_vals_fine"])], axis=1), axis=1))
    k = np.arange(1, Nf + 1)
    mean_k = k / (Nf + 1.0)
    var_k = k * (Nf + 1.0 - k) / ((Nf + 1.0) ** 2 * (Nf + 2.0))
    # mean of every order statistic: standard error sqrt(var_k / B) <= 5e-4; 6 sigma
    assert np.abs(u.mean(0) - mean_k).max() <= 6 * np.sqrt(var_k.max() / B) + 2e-6
    # variance of every order statistic within 12 % (relative s.e. of a variance estimate ~ sqrt(2/B) <= 2.2 %)
    assert np.abs(u.var(0) / var_k - 1).max() <= 0.12
    # a long gap U_(3Nf/4) - U_(Nf/4) is Beta(Nf/2, Nf/2+1): tests the joint law, not only the marginals
    gap = u[:, 3 * Nf // 4 - 1] - u[:, Nf // 4 - 1]
    a, b = Nf / 2.0, Nf / 2.0 + 1.0
    assert abs(gap.mean() - a / (a + b)) <= 1e-3 and abs(gap.var() / (a * b / ((a + b) ** 2 * (a + b + 1))) - 1) <= 0.12
    # pooled, the samples are uniform
    assert abs(u.mean() - 0.5) <= 1e-3 and abs(u.var() - 1 / 12) <= 1e-3
    # different rays and different seeds give different draws; the same seed repeats
    assert not np.array_equal(u[0], u[1])
    ts2 = ru.sample_fine(Nf, dev(w), dev(edges), dev(tc), None, seed=22)
    assert not torch.equal(ts2, ts) and torch.equal(ru.sample_fine(Nf, dev(w), dev(edges), dev(tc), None, seed=21), ts)


@pytest.mark.parametrize("Nc", [2, 7, 10, 64, 130])
@pytest.mark.parametrize("lin_inv", [True, False])
def test_coarse_sampler_any_bin_count_bit_exact(Nc, lin_inv):
    """The stratified sampler works on groups of 4 bins per thread; bin counts that are not multiples of 4, with
    and without perturbation, must still be bit-identical to the oracle."""
    rng = np.random.default_rng(100 + Nc)
    B = 53
    near = (rng.random((B, 1), dtype=F32) * F32(0.3) + F32(0.3)).astype(F32)
    far = (near + F32(0.5) + rng.random((B, 1), dtype=F32)).astype(F32)
    u = rng.random((B, Nc), dtype=F32)
    z = np.zeros((B, 3), F32)
    for perturb, uu in ((True, u), (False, None)):
        o = rm.create_input_batch_coarse_model(Nc, lin_inv, perturb, z, z + 1, near, far, uu)
        t, e = ru.sample_coarse(Nc, lin_inv, perturb, dev(near), dev(far), None if uu is None else dev(uu))
        assert np.array_equal(host(e), o["bin_data"]["bin_edges"])
        assert np.array_equal(host(t), o["t_vals"])
