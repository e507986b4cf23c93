"""GPU tests of the callers either side of the hot path: on-device sample-mode input,
per-view render driver / post-processing, checkpoint round trip; plus the cfg5 sampling shape and
size-independent properties at BASELINE.json's full 800x800 size."""
import os

import numpy as np
import pytest
import torch

import nerf_tf2_b200 as nb
from nerf_tf2_b200 import ray_utils as ru
from oracle import model as om, ray_march as rm, scene as osc

pytestmark = pytest.mark.gpu
F32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def test_sample_mode_dataset_matches_reference_map_function():
    """_sample_mode_map_function (core/base_dataset.py:555-621): all rays (get_rays_tf) + gather by idxs."""
    rng = np.random.default_rng(0)
    H, W, N = 20, 30, 3
    imgs = rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)
    views = [osc.synthetic_view(H, W, view=i) for i in range(N)]
    ds = nb.SampleModeDataset(imgs, [v["c2w"] for v in views], views[0]["bounds"], views[0]["K"], batch_size=512, seed=7)
    for img_i, step in ((0, 0), (2, 5)):
        ((ro, rd, near, far), (rgb,)), ids = ds.draw(img_i, step)
        ids_h = host(ids)
        assert ids_h.dtype == np.int32 and ids_h.min() >= 0 and ids_h.max() < H * W
        all_o, all_d = rm.get_rays_f32(H, W, views[img_i]["K"], views[img_i]["c2w"])
        assert np.abs(host(rd) - all_d[ids_h]).max() <= 1.2e-7 and np.array_equal(host(ro), all_o[ids_h])
        ref_rgb = (imgs[img_i].reshape(-1, 3).astype(F32) / F32(255.0))[ids_h]
        assert np.array_equal(host(rgb), ref_rgb)                     # byte work: bit-exact
        assert np.all(host(near) == views[0]["near"][0, 0]) and near.shape == (512, 1)
    a = host(ds.draw(0, 1)[1]); b = host(ds.draw(0, 2)[1])
    assert not np.array_equal(a, b) and np.array_equal(a, host(ds.draw(0, 1)[1]))
    big = nb.SampleModeDataset(imgs, [v["c2w"] for v in views], views[0]["bounds"], views[0]["K"], batch_size=60000, seed=1)
    hist = np.bincount(host(big.draw(1, 0)[1]), minlength=H * W)
    assert hist.min() > 50 and hist.max() < 160                       # uniform over the 600 pixels (mean 100)
    it = iter(ds)
    batch = next(it)
    assert len(batch) == 2 and batch[0][0].shape == (512, 3) and batch[1][0].shape == (512, 3)


def test_render_view_postprocessing_matches_reference_scripts():
    H, W = 24, 20
    v = osc.synthetic_view(H, W, view=3)
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=False), precision="fp32", seed=5)
    gt = np.random.default_rng(1).integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    out = nb.render.render_view(nerf, H, W, v["c2w"], v["bounds"], v["K"], gt_u8=gt, scale_factor=v["adj_scale_factor"])
    pred = host(out["pred_rgb"]).astype(F32)
    # main/render.py:96-97 / main/eval.py:53-60
    clipped = np.clip(pred * F32(255.0), F32(0.0), F32(255.0))
    assert np.array_equal(host(out["img_u8"]), clipped.astype(np.uint8))
    ref_psnr = rm.psnr_metric_numpy(gt.reshape(-1, 3).astype(F32) / 255.0, clipped.astype(F32) / 255.0)
    assert abs(out["psnr"] - float(ref_psnr)) <= 1e-4
    depth = host(out["pred_depth"])
    # main/render.py:103-112: both depth maps are taken with the camera->W2 pose (the spherical-path pose BEFORE the
    # scene scale); render_view undoes the scale of the W3 pose it was given
    c2w_W2 = rm.create_spherical_path(4.0, 40.0, 8)[3]
    for mt in ("type_1", "type_2"):
        ref = rm.create_depth_map(depth, H, W, v["adj_scale_factor"], mt, v["K"], c2w_W2).reshape(-1)
        assert np.allclose(host(out[f"depth_{mt}"]), ref, rtol=2e-6, atol=2e-6)
    # type_2 is a camera-space z: positive, and never larger than the type_1 distance along the ray
    z, dist = host(out["depth_type_2"]), host(out["depth_type_1"])
    assert np.all(z > 0) and np.all(z <= dist * (1 + 1e-5))
    # a ray sub-range renders the same pixels as the full view (fixed sampling: perturb off would still
    # draw random fine uniforms, keyed by GLOBAL ray id, so ranges agree exactly)
    part = nb.render.render_view(nerf, H, W, v["c2w"], v["bounds"], v["K"], ray0=100, n_rays=160)
    assert torch.equal(part["img_u8"], out["img_u8"][100:260])


def test_evaluate_views_mean_psnr():
    H = W = 12
    views = [osc.synthetic_view(H, W, view=i) for i in range(3)]
    gts = np.random.default_rng(2).integers(0, 256, size=(3, H, W, 3), dtype=np.uint8)
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}), precision="bf16", seed=2)
    res = nb.render.evaluate_views(nerf, H, W, [v["c2w"] for v in views], views[0]["bounds"], views[0]["K"], gts)
    assert res["psnr_vals"].shape == (3,) and np.all(np.isfinite(res["psnr_vals"]))
    assert abs(res["mean_psnr"] - res["psnr_vals"].mean()) < 1e-12 and res["last_view_psnr"] == res["psnr_vals"][-1]


def test_checkpoint_round_trip(tmp_path):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_forward_train.npz"))
    batch = ((g["rays_o"], g["rays_d"], g["near"], g["far"]), (g["rgb_gt"],))
    a = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=False), precision="bf16", seed=4)
    for _ in range(3):
        a.train_step(batch, u_fine=dev(g["u_fine"]))
    saver = nb.CustomSaver(str(tmp_path), save_best_only=False)
    saver.set_model(a)
    saver.on_epoch_end(7, {"psnr_metric": 11.5})                                   # no validation -> nothing saved
    assert not any(f.endswith(".npz") for f in os.listdir(tmp_path))
    saver.on_epoch_end(8, {"psnr_metric": 11.75, "val_psnr_metric": 12.3456})
    tag = "000008_12.35"
    assert sorted(os.listdir(tmp_path)) == [f"{tag}_coarse.npz", f"{tag}_fine.npz", f"{tag}_logs.npz", f"{tag}_optimizer.npz"]
    opt = np.load(tmp_path / f"{tag}_optimizer.npz")
    names = [str(n) for n in opt["names"]]
    assert len(names) == 97 and names[0] == "Adam/iter:0" and names[1] == "Adam/coarse/dense_0/kernel/m:0"
    assert int(opt["Adam/iter:0"]) == 3 and opt[names[1]].shape == (63, 256)
    b = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=False), precision="bf16", seed=99)
    b.set_everything(str(tmp_path), tag)
    assert torch.equal(a.flat_params, b.flat_params) and torch.equal(a.optimizer.m, b.optimizer.m)
    assert torch.equal(a.optimizer.v, b.optimizer.v) and b.optimizer.iterations == 3
    b._step_counter = a._step_counter
    a.train_step(batch, u_fine=dev(g["u_fine"])); b.train_step(batch, u_fine=dev(g["u_fine"]))
    assert torch.allclose(a.flat_params, b.flat_params, atol=1e-6)                 # resumed run continues identically
    logs = np.load(tmp_path / f"{tag}_logs.npz")
    assert list(logs["val_epoch_idxs"]) == [8] and list(logs["train_epoch_idxs"]) == [7, 8]


def test_cfg5_sampling_shape_128_256():
    """BASELINE config 5 shape: 128 coarse + 256 fine samples (S = 384 rows per ray in the fine pass)."""
    v = osc.synthetic_view(6, 6, view=1)
    rng = np.random.default_rng(3)
    uf = rng.random((36, 256), dtype=F32)
    w = om.init_weights(9)
    pc, pf = om.forward(w, v["rays_o"], v["rays_d"], v["near"], v["far"], N_coarse=128, N_fine=256, u_fine=uf,
                        perturb=False, white_bg=False)
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": False}}, N_coarse=128, N_fine=256, perturb=False), precision="fp32")
    nerf.set_weights_from_dict(w)
    oc, of = nerf.forward(dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]), u_fine=dev(uf))
    assert of["weights"].shape == (36, 384) and oc["weights"].shape == (36, 128)
    assert np.abs(host(of["pred_rgb"]) - pf["pred_rgb"]).max() <= 5e-4
    assert np.abs(host(of["pred_depth"]) - pf["pred_depth"]).max() <= 1e-3
    nerf16 = nb.setup_model(nb.make_params({"system": {"white_bg": False}}, N_coarse=128, N_fine=256, perturb=False), precision="bf16")
    nerf16.set_weights_from_dict(w)
    _, of16 = nerf16.forward(dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]), u_fine=dev(uf))
    assert np.percentile(np.abs(host(of16["pred_rgb"]) - pf["pred_rgb"]), 95) <= 1.5e-2


def test_full_size_800x800_properties():
    """Size-independent properties at BASELINE.json's full size (640 000 rays, 64+128 samples)."""
    H = W = 800
    torch.manual_seed(20261017)                                       # the synthetic weights below: reproducible
    sc = nb.scene.SyntheticScene(H, W)
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}), precision="bf16", seed=0)
    ds = nb.create_dataset_for_render(H, W, sc.poses[2], sc.bounds, sc.K, on_device=True)
    ro, rd, near, far = ds.inputs
    n = H * W
    assert len(ds) == 157                                             # ceil(640000 / 4096) reference chunks
    # samplers: stratified t inside its bin; hierarchical output ascending and inside [near, far] (+ulps)
    t_c, edges = ru.sample_coarse(64, True, True, near, far, None, seed=3)
    assert bool(((t_c >= edges[:, :-1]) & (t_c <= edges[:, 1:])).all())
    wts = torch.rand((n, 64), device="cuda") ** 6
    t_f = ru.sample_fine(128, wts, edges, t_c, None, seed=3)
    assert t_f.shape == (n, 192) and bool((t_f[:, 1:] >= t_f[:, :-1]).all())
    # (u - cdf) / pdf amplifies the fp32 rounding of the CDF where the pdf is tiny (weights ~ rand^6), in the reference's
    # formula as here, so a sample may overshoot its bin by a fraction of the bin width (0.85 / 64)
    assert float(t_f.min()) >= sc.near - 1e-6 and float(t_f.max()) <= sc.far + 2e-3
    # coarse samples survive the merge: every t_c value appears in the sorted output
    assert bool((torch.searchsorted(t_f, t_c) < 192).all())
    # integrator: acc = sum(weights) in [0, 1], white background keeps rgb in [0, 1]
    sig = torch.rand((n * 64,), device="cuda") * 30 * (torch.rand((n * 64,), device="cuda") > 0.6)
    rgb = torch.rand((n * 64, 3), device="cuda")
    pp = ru.post_process_model_output(rgb, sig, t_c, True)
    assert torch.allclose(pp["weights"].sum(1), pp["acc_map"], atol=2e-6)
    assert float(pp["acc_map"].min()) >= 0 and float(pp["acc_map"].max()) <= 1 + 1e-5
    assert float(pp["pred_rgb"].min()) >= -1e-5 and float(pp["pred_rgb"].max()) <= 1 + 1e-5
    # whole pipeline: finite, in range, and independent of the chunking (Philox keyed by global ray id)
    out = nb.render.render_view(nerf, H, W, sc.poses[2], sc.bounds, sc.K, scale_factor=sc.adj_scale_factor)
    assert out["img_u8"].shape == (n, 3) and bool(torch.isfinite(out["pred_rgb"]).all())
    assert float(out["acc_map"].min()) >= 0 and float(out["acc_map"].max()) <= 1 + 1e-5
    nerf.render_chunk = 50000
    out2 = nb.render.render_view(nerf, H, W, sc.poses[2], sc.bounds, sc.K, depth_maps=False)
    assert torch.equal(out["img_u8"], out2["img_u8"])


def test_bench_line_contract():
    """One short bench run on the GPU: ONE JSON line with the contract's keys (value/e2e/roofline/gpu_launches/clocks)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu", "--no-train", "--no-extra"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["unit"] == "rays/s" and d["n_gpus"] == 1 and d["scaling"] == "strong" and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert d["value"] > 1e5 and d["gpu_launches"] > 0 and d["warmup"] >= 3
    r = d["roofline"]
    assert r["bound"] == "tensor" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # traffic comes from the committed ncu export the line names (never a hand-maintained constant)
    assert r["traffic"] > 0 and r["traffic_source"].startswith("profiles/") and os.path.exists(os.path.join(root, r["traffic_source"].split(":")[0]))
    e = d["e2e"]
    assert e["value"] > 1e5 and e["h2d_bytes_per_step"] == 640000 * 8 * 4 and e["d2h_bytes_per_step"] == 640000 * 10 * 4
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"] and "workload" in d["config"]
    # the HBM-bound kernels, timed alone against the measured copy bandwidth
    hb = d["roofline_hbm"]
    assert len(hb) == 3 and all(h["bound"] == "hbm" and 0 < h["frac"] < 1.2 and h["in_step_GBps"] > 0 for h in hb)
    assert all(abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-9 for h in hb)


def test_render_spherical_path_script_loop(tmp_path):
    """main/render.py:51-117 as a function: poses on a sphere, W3 scaling, uint8 frames + both depth maps + acc map,
    files named like the script's."""
    params = nb.make_params({"system": {"white_bg": True}, "render": {"radius": 4.0, "inclination": 40.0, "num_cameras": 3,
                                                                      "img_size": [10, 12], "camera_model_name": "SIMPLE_PINHOLE",
                                                                      "camera_model_params": [16.0, 6.0, 5.0], "bounds": None}},
                            perturb=False)
    nerf = nb.setup_model(params, precision="bf16", seed=3)
    frames = nb.render.render_spherical_path(nerf, params.render, adj_scale_factor=0.2125, save_dir=str(tmp_path))
    assert len(frames) == 3 and frames[0]["img_u8"].shape == (10, 12, 3) and frames[0]["img_u8"].dtype == np.uint8
    assert frames[1]["depth_type_1"].shape == (10, 12) and frames[1]["depth_type_2"].shape == (10, 12)
    assert sorted(os.listdir(tmp_path / "rgb")) == ["render_00000.png", "render_00001.png", "render_00002.png"]
    assert np.array_equal(np.load(tmp_path / "acc_map" / "render_00002.npy"), frames[2]["acc_map"])
    # the same view through render_view with explicit W3 pose / bounds (0.25 and 0.75 of the diameter, scaled)
    from nerf_tf2_b200 import pose_utils as pu
    pose2 = pu.create_spherical_path(4.0, 40.0, 3, None)[1]
    pose3, b3 = pu.reconfigure_scene_scale(pose2, np.array([2.0, 6.0]), 0.2125)
    K = nb.CustomDataset.camera_model_params_to_intrinsics("SIMPLE_PINHOLE", [16.0, 6.0, 5.0])
    r = nb.render.render_view(nerf, 10, 12, pose3, b3, K, scale_factor=0.2125, c2w_W2=pose2)
    assert np.array_equal(host(r["img_u8"]).reshape(10, 12, 3), frames[1]["img_u8"])


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_one_call_forward_equals_step_by_step(precision):
    """nerfb200_forward (the whole march of NeRF.forward, core/model.py:57-125, as one C-ABI call) returns bit for bit
    what the same launches issued one by one from Python return: fixed uniforms and in-kernel Philox, with and without
    the per-sample weights."""
    import numpy as np
    from oracle import scene as osc
    v = osc.synthetic_view(37, 29, view=3)
    n = 37 * 29
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ro, rd, near, far = (dv(v[k]) for k in ("rays_o", "rays_d", "near", "far"))
    uf = torch.rand((n, 128), device="cuda")
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=True), precision=precision, seed=5, rng_seed=11)
    for kw in (dict(u_fine=uf), dict(), dict(need_weights=False), dict(ray0=12345)):
        nerf.fused_forward = True
        c1, f1 = nerf.forward(ro, rd, near, far, **kw)
        nerf.fused_forward = False
        c0, f0 = nerf.forward(ro, rd, near, far, **kw)
        for a, b in ((c1, c0), (f1, f0)):
            keys = [k for k in b if k in a]
            assert set(keys) >= {"pred_rgb", "pred_depth", "acc_map"} and ("weights" in a) == (kw.get("need_weights", True))
            for k in keys:
                assert torch.equal(a[k], b[k]), (precision, kw.keys(), k)
    lib = nb._lib.load()
    assert lib.nerfb200_forward_workspace_bytes(0, 64, 128) == 0 and lib.nerfb200_forward_workspace_bytes(-1, 64, 128) == -1
