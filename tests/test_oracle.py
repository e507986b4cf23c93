"""CPU tests of the oracle itself: against the committed golden fixtures (produced by the
reference's own ray_utils.py over the TF op-shim) and, where /root/reference exists, against
that code executed live."""
import numpy as np
import pytest

from conftest import HAVE_REFERENCE
from oracle import model, ray_march as rm, scene

F32 = np.float32


def test_rays_match_reference_fixture(golden):
    g = golden["ref_rays"]
    for tag in ("a", "b"):
        H, W = int(g[f"{tag}_H"]), int(g[f"{tag}_W"])
        ro, rd = rm.get_rays(H, W, g[f"{tag}_K"], g[f"{tag}_c2w"])
        assert np.array_equal(ro.astype(F32), g[f"{tag}_rays_o"])
        assert np.array_equal(rd.astype(F32), g[f"{tag}_rays_d"])
        ro32, rd32 = rm.get_rays_f32(H, W, g[f"{tag}_K"], g[f"{tag}_c2w"])
        assert np.array_equal(ro32, g[f"{tag}_rays_o_tf"])
        assert np.array_equal(rd32, g[f"{tag}_rays_d_tf"])
    assert np.array_equal(rm.create_spherical_path(4.0, 40.0, 8), g["spherical_path_r4_i40_n8"])


@pytest.mark.parametrize("tag,lin_inv", [("inv", True), ("lin", False)])
def test_sampling_and_compositing_match_reference_fixture(golden, tag, lin_inv):
    g = golden["ref_sampling_composite"]
    d = rm.create_input_batch_coarse_model(64, lin_inv, True, g["rays_o"], g["rays_d"], g["near"], g["far"],
                                           g[f"{tag}_u_coarse"])
    assert np.array_equal(d["t_vals"], g[f"{tag}_t_coarse"])
    assert np.array_equal(d["bin_data"]["bin_edges"], g[f"{tag}_bin_edges"])
    assert np.array_equal(d["xyz_inputs"], g[f"{tag}_xyz_coarse"])
    for wb in (0, 1):
        pp = rm.post_process_model_output(g[f"{tag}_rgb"], g[f"{tag}_sigma"], d["t_vals"], bool(wb))
        for k in ("acc_map", "weights", "pred_rgb", "pred_depth"):
            assert np.array_equal(pp[k], g[f"{tag}_wb{wb}_{k}"]), k
    f = rm.create_input_batch_fine_model(g["rays_o"], g["rays_d"], g[f"{tag}_wb0_weights"], d["bin_data"],
                                         d["t_vals"], g[f"{tag}_u_fine"], return_debug=True)
    assert np.array_equal(f["t_vals"], g[f"{tag}_t_fine_sorted"])
    assert np.array_equal(f["piece_idxs"], g[f"{tag}_piece_idxs"])
    assert f["piece_idxs"].dtype == np.int32 and f["piece_idxs"].min() >= 0 and f["piece_idxs"].max() <= 63
    assert np.all(np.diff(f["t_vals"], axis=1) >= 0)


def test_perturb_off_is_midpoints():
    v = scene.synthetic_view(4, 4)
    d = rm.create_input_batch_coarse_model(64, True, False, v["rays_o"], v["rays_d"], v["near"], v["far"])
    e = d["bin_data"]["bin_edges"]
    assert np.array_equal(d["t_vals"], F32(0.5) * (e[:, :-1] + e[:, 1:]))
    assert np.array_equal(e[:, 0], v["near"][:, 0]) and np.allclose(e[:, -1], v["far"][:, 0], rtol=1e-6)


def test_sampler_edge_cases():
    B = 4
    edges = rm.tf_linspace(np.full((B, 1), 0.4, F32), np.full((B, 1), 1.2, F32), 65)
    bd = {"left_edges": edges[:, :-1], "bin_widths": edges[:, 1:] - edges[:, :-1]}
    tc = F32(0.5) * (edges[:, :-1] + edges[:, 1:])
    w = np.zeros((B, 64), F32)
    w[1, 10] = 1.0                                   # one-hot
    w[2] = 1.0
    u = np.random.default_rng(0).random((B, 128), dtype=F32)
    u[0, 0], u[0, 1] = 0.0, np.nextafter(F32(1), F32(0))
    ro = np.zeros((B, 3), F32)
    rd = np.tile(np.array([[0, 0, 1]], F32), (B, 1))
    out = rm.create_input_batch_fine_model(ro, rd, w, bd, tc, u, return_debug=True)
    assert out["t_vals"].shape == (B, 192) and np.all(np.diff(out["t_vals"], axis=1) >= 0)
    # all-zero weights -> uniform pdf -> fine samples uniform over [near, far]
    assert abs(out["t_vals_fine"][0].mean() - 0.8) < 0.05
    # one-hot weights -> almost all fine samples inside bin 10
    inside = (out["piece_idxs"][1] == 10).mean()
    assert inside > 0.95
    assert out["cdf"][:, 0].max() == 0 and np.all(np.abs(out["cdf"][:, -1] - 1) < 1e-5)


def test_mlp_and_forward_fixture_regenerates(golden):
    g = golden["oracle_mlp"]
    w = model.init_weights(int(g["weights_seed"]), bias_scale=float(g["bias_scale"]))
    import torch
    for m in ("coarse", "fine"):
        rgb, sig = model.mlp_forward_np(w, m, g["xyz"], g["dirs"], torch.float32)
        # same code, same inputs: must regenerate up to BLAS thread-count reassociation
        assert np.allclose(rgb, g[f"{m}_rgb_f32"], atol=2e-5)
        assert np.allclose(sig, g[f"{m}_sigma_f32"], atol=2e-4, rtol=1e-4)
        # fp32 vs fp64 noise floor of the network itself
        assert np.abs(rgb - g[f"{m}_rgb_f64"]).max() < 5e-4
    enc = model.positional_encode(torch.from_numpy(g["xyz"]), 10).numpy()
    assert enc.shape == (g["xyz"].shape[0], 63)
    assert np.array_equal(enc[:, :3], g["xyz"])
    # layout j = 3 + d*2L + 2l + s  (core/model.py:325-330)
    x = g["xyz"].astype(np.float64)
    m9 = float(F32(512.0) * F32(np.pi))
    assert np.allclose(enc[:, 3 + 1 * 20 + 2 * 9 + 1], np.cos(np.float32(g["xyz"][:, 1] * F32(m9)).astype(np.float64)), atol=2e-6)


def test_psnr_definitions_differ_by_log10_3():
    rng = np.random.default_rng(0)
    y, p = rng.random((100, 3), dtype=F32), rng.random((100, 3), dtype=F32)
    m = rm.PSNRMetric()
    m.update_state(y[:50], p[:50]); m.update_state(y[50:], p[50:])
    assert abs((rm.psnr_metric_numpy(y, p) - m.result()) - 10 * np.log10(3)) < 1e-4


def test_adam_matches_closed_form_first_step():
    w = {"a": np.array([1.0, -2.0], F32)}
    g = {"a": np.array([0.5, -0.25], F32)}
    m = {"a": np.zeros(2, F32)}; v = {"a": np.zeros(2, F32)}
    it = model.adam_step(w, g, m, v, 0)
    # first Adam step moves each weight by ~lr * sign(g)
    assert it == 1 and np.allclose(w["a"], [1.0 - 5e-4, -2.0 + 5e-4], atol=1e-7)
    assert abs(model.exponential_decay_lr(500000) - 5e-5) < 1e-12


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference not present (GPU box)")
def test_oracle_bit_identical_to_reference_live():
    from oracle import tf_shim
    rng = np.random.default_rng(3)
    v = scene.synthetic_view(9, 7, view=4)
    B = 63
    with tf_shim.reference_ray_utils() as (ru, pu, tf):
        ro, rd = ru.get_rays(9, 7, v["K"], v["c2w"])
        assert np.array_equal(ro.astype(F32), v["rays_o"]) and np.array_equal(rd.astype(F32), v["rays_d"])
        for lin_inv in (True, False):
            uc, uf = rng.random((B, 64), dtype=F32), rng.random((B, 128), dtype=F32)
            tf.random.queue = [uc, uf]
            p = tf_shim.Params(perturb=True, lin_inv_depth=lin_inv)
            d = ru.create_input_batch_coarse_model(p, v["rays_o"], v["rays_d"], v["near"], v["far"])
            mine = rm.create_input_batch_coarse_model(64, lin_inv, True, v["rays_o"], v["rays_d"], v["near"],
                                                      v["far"], uc)
            assert np.array_equal(d["t_vals"], mine["t_vals"]) and np.array_equal(d["xyz_inputs"], mine["xyz_inputs"])
            sig = (rng.random((B * 64, 1), dtype=F32) * 30 * (rng.random((B * 64, 1)) > 0.5)).astype(F32)
            rgb = rng.random((B * 64, 3), dtype=F32)
            pp = ru.post_process_model_output(rgb, sig, d["t_vals"], True)
            pm = rm.post_process_model_output(rgb, sig, mine["t_vals"], True)
            for k in pp:
                assert np.array_equal(pp[k], pm[k]), k
            f = ru.create_input_batch_fine_model(p, v["rays_o"], v["rays_d"], pp["weights"], d["bin_data"], d["t_vals"])
            fm = rm.create_input_batch_fine_model(v["rays_o"], v["rays_d"], pm["weights"], mine["bin_data"],
                                                  mine["t_vals"], uf)
            for k in f:
                assert np.array_equal(f[k], fm[k]), k
        with pytest.raises(UnboundLocalError):   # the reference cannot run perturb=False
            ru.create_input_batch_coarse_model(tf_shim.Params(perturb=False), v["rays_o"], v["rays_d"], v["near"], v["far"])
