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
