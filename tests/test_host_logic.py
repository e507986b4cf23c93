"""CPU: host-side logic of the package (dataset batching, params, PSNR, init, scene)."""
import numpy as np
import pytest
import torch

import nerf_tf2_b200 as nb
from nerf_tf2_b200 import model as pm, scene as ps
from oracle import model as om, ray_march as rm, scene as osc


def test_glorot_init_matches_oracle_init():
    flat = pm.glorot_uniform_params(7)
    w = om.init_weights(7)
    cat = np.concatenate([w[n].reshape(-1) for n in om.all_variable_names()])
    assert np.array_equal(flat, cat)
    assert pm.variable_names("coarse") == om.variable_names("coarse")


def test_scene_matches_oracle_scene():
    sc = ps.SyntheticScene(10, 12, num_cameras=8)
    for i in (0, 3, 7):
        v = osc.synthetic_view(10, 12, view=i)
        assert np.allclose(sc.poses[i], v["c2w"], atol=1e-12)
        assert np.array_equal(sc.K, v["K"])
        assert np.float32(sc.near) == v["near"][0, 0] and np.float32(sc.far) == v["far"][0, 0]
    assert np.allclose(ps.spherical_poses(4.0, 40.0, 8), rm.create_spherical_path(4.0, 40.0, 8), atol=1e-12)


def test_ray_dataset_batching_semantics():
    n = 10
    a = [np.arange(n * k, dtype=np.float32).reshape(n, k) for k in (3, 3, 1, 1)]
    ds = nb.RayDataset.from_tensor_slices((tuple(a),)).batch(4, drop_remainder=False)
    batches = list(ds)
    assert [b[0][0].shape[0] for b in batches] == [4, 4, 2] and len(ds) == 3
    assert len(batches[0]) == 1 and len(batches[0][0]) == 4          # ((ro, rd, near, far),)
    ds2 = nb.RayDataset.from_tensor_slices((tuple(a), (a[0],))).batch(4, drop_remainder=True)
    b2 = list(ds2)
    assert len(b2) == 2 and len(b2[0]) == 2 and b2[0][1][0].shape == (4, 3)
    # repeat + shuffle: every epoch is a permutation, batches stay full
    ds3 = nb.RayDataset.from_tensor_slices((tuple(a), (a[0],))).shuffle(seed=1).repeat().batch(5, drop_remainder=True)
    it = iter(ds3)
    seen = np.concatenate([next(it)[0][2][:, 0] for _ in range(2)])
    assert sorted(seen.tolist()) == list(range(n))


def test_params_defaults_and_overrides():
    p = nb.make_params({"system": {"white_bg": True}}, N_fine=256)
    assert p.system.white_bg is True and p.sampling.N_fine == 256 and p.sampling.N_coarse == 64
    assert p.data.batch_size == 4096 and p.sampling.lin_inv_depth is True


def test_psnr_functions():
    rng = np.random.default_rng(0)
    y, p = rng.random((64, 3), dtype=np.float32), rng.random((64, 3), dtype=np.float32)
    m = nb.PSNRMetric()
    m.update_state(torch.from_numpy(y[:32]), torch.from_numpy(p[:32]))
    m.update_state(torch.from_numpy(y[32:]), torch.from_numpy(p[32:]))
    o = rm.PSNRMetric(); o.update_state(y, p)
    assert abs(m.result() - float(o.result())) < 1e-4
    assert abs(nb.psnr_metric_numpy(y, p) - rm.psnr_metric_numpy(y, p)) < 1e-12
    assert abs(float(nb.psnr_metric(torch.from_numpy(y), torch.from_numpy(p))) - rm.psnr_metric_numpy(y, p)) < 1e-4
    m.reset_states()
    assert m.state.sum() == 0


def test_checkpoint_file_format(tmp_path):
    """CustomSaver._save_weights format (core/ops.py:110-120): name -> array plus an ordered `names` array."""
    from nerf_tf2_b200 import checkpoint

    class V:
        def __init__(self, name, a):
            self.name, self._a = name, a

        def numpy(self):
            return self._a
    vs = [V("coarse/dense_0/kernel", np.ones((63, 256), np.float32)), V("coarse/dense_0/bias", np.zeros(256, np.float32))]
    path = tmp_path / "w.npz"
    checkpoint.save_weights(str(path), vs)
    d = np.load(path)
    assert [str(n) for n in d["names"]] == ["coarse/dense_0/kernel", "coarse/dense_0/bias"]
    assert d["coarse/dense_0/kernel"].shape == (63, 256)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port) must print exactly ONE JSON line with
    the contract's keys; runs on CPU in a few seconds."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_metric_value_is_a_lazy_number():
    """train_step / test_step hand back PSNRMetric.result_async(): a snapshot that turns into a float on first use."""
    import json
    import torch
    m = nb.PSNRMetric()
    y = torch.rand(16, 3); p = torch.rand(16, 3)
    m.update_state(y, p)
    v, ref = m.result_async(), m.result()
    m.update_state(y, p * 0.5)                      # later updates do not change the snapshot
    assert float(v) == ref and abs(v - ref) == 0 and v == ref and not (v < ref) and v >= ref
    assert np.isfinite(v) and f"{v:.3f}" == f"{ref:.3f}" and repr(v) == repr(ref)
    assert (v + 1) - 1 == pytest.approx(ref) and 2 * v == pytest.approx(2 * ref) and -v == -ref
    assert json.dumps({"psnr": float(v)}) == json.dumps({"psnr": ref})
    assert m.result() != ref


def test_render_tolerance_statement():
    """oracle/tolerance.py: p99 + bounded maximum over every pixel; only rays whose last-sample alpha flipped are exempt,
    and only up to the stated rate."""
    from oracle.tolerance import MAX_FLIP_RATE, RENDER_TOL, check_render, last_alpha_flips
    rng = np.random.default_rng(0)
    n = 4000

    def outputs(S):
        w = rng.random((n, S), dtype=np.float32) * 0.01 + 1e-3
        return {"pred_rgb": rng.random((n, 3), dtype=np.float32), "pred_depth": rng.random(n, dtype=np.float32),
                "acc_map": rng.random(n, dtype=np.float32), "weights": w}
    ref_c, ref_f = outputs(64), outputs(192)
    copy = lambda d: {k: v.copy() for k, v in d.items()}
    g_c, g_f = copy(ref_c), copy(ref_f)
    g_f["pred_rgb"] += 1e-3
    m = check_render("bf16", g_c, g_f, ref_c, ref_f)
    assert m["last_alpha_flips"] == 0 and abs(m["fine_pred_rgb"]["max_over_all"] - 1e-3) < 1e-6
    # one pixel beyond the maximum, last-sample alpha unchanged: not exempt
    bad = copy(g_f); bad["pred_rgb"][7, 1] += 0.3
    with pytest.raises(AssertionError):
        check_render("bf16", g_c, bad, ref_c, ref_f)
    # the same pixel with a flipped last-sample alpha (w_last 0 <-> >0) is the counted exception ...
    flip = copy(bad); flip["weights"][7, -1] = 0.0
    assert last_alpha_flips(g_c, flip, ref_c, ref_f).sum() == 1
    assert check_render("bf16", g_c, flip, ref_c, ref_f)["last_alpha_flips"] == 1
    # ... but only up to the stated rate
    many = copy(flip); many["weights"][:10, -1] = 0.0
    assert MAX_FLIP_RATE * n < 10
    with pytest.raises(AssertionError):
        check_render("bf16", g_c, many, ref_c, ref_f)
    assert RENDER_TOL["fp16"]["pred_rgb"][1] < RENDER_TOL["bf16"]["pred_rgb"][1]


def test_bench_reads_traffic_from_the_committed_ncu_exports():
    """roofline.traffic is parsed from the `ncu --page raw --csv` exports under profiles/ that the line names."""
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.ncu_traffic(bench.NCU_MLP_CSV["bf16"], "mlp_tc_forward_pair_kernel<0, 0, 0, 0>")
    assert src.startswith("profiles/") and 1.5e8 < t < 3.0e8          # ~200 MB per fine launch (50 MB read + 147 MB written)
    t32, _ = bench.ncu_traffic(bench.NCU_MLP_CSV["tf32"], "mlp_tf32_forward_kernel")
    assert 1.5e8 < t32 < 3.0e8
    th, _ = bench.ncu_traffic(bench.NCU_HBM_CSV, "sample_fine_fast_kernel")
    assert 3e7 < th < 1.2e8
    assert bench.ncu_traffic("no_such_file.csv", "x") == (None, None)
    assert bench.ncu_traffic(bench.NCU_HBM_CSV, "no_such_kernel") == (None, None)


def test_documents_only_cite_files_that_exist():
    """DESIGN.md / README.md / INTEGRATION.md / tools/README.md cite measurements and sources by path: every cited path
    under profiles/, tools/, tests/, oracle/, include/ and the package must exist in the tree."""
    import os
    import re
    from conftest import ROOT
    missing = []
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("tools", "README.md")):
        txt = open(os.path.join(ROOT, doc)).read()
        for m in re.findall(r"((?:profiles|tools|tests|oracle|include|nerf-tf2_b200)/[A-Za-z0-9_\-./*{},<>…]+)", txt):
            m = m.rstrip(".,)`:;")
            if any(c in m for c in "*{}<>…") or m == "oracle/_ref":      # patterns; the directory that by design does not exist
                continue
            if not os.path.exists(os.path.join(ROOT, m)):
                missing.append((doc, m))
    assert not missing, missing
