"""
TEST INFRASTRUCTURE ONLY -- generates the committed fixtures under tests/golden/.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden

Fixtures whose name starts with `ref_` are outputs of the reference's OWN
`nerf/utils/ray_utils.py` / `pose_utils.py`, imported unmodified from /root/reference and
executed over oracle/tf_shim.py with `tf.random.uniform` replaced by fixed uniforms
(perturb=True path, the only one that runs in the reference).
Fixtures starting with `oracle_` come from oracle/model.py (TensorFlow/Keras are not
installable, so the MLP/Adam fixtures are restatement outputs: parity unpinned).
"""
import os

import numpy as np

from . import model, ray_march as rm, scene, tf_shim

F32 = np.float32
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen_rays(ru, pu):
    out = {}
    for tag, (H, W, view) in {"a": (12, 10, 1), "b": (7, 9, 5)}.items():
        v = scene.synthetic_view(H, W, view=view)
        ro, rd = ru.get_rays(H, W, v["K"], v["c2w"])
        ro32, rd32 = ru.get_rays_tf(H, W, v["K"].astype(F32), v["c2w"].astype(F32))
        out.update({f"{tag}_H": H, f"{tag}_W": W, f"{tag}_K": v["K"], f"{tag}_c2w": v["c2w"],
                    f"{tag}_rays_o": ro.astype(F32), f"{tag}_rays_d": rd.astype(F32),
                    f"{tag}_rays_o_tf": np.ascontiguousarray(ro32), f"{tag}_rays_d_tf": rd32})
    out["spherical_path_r4_i40_n8"] = pu.create_spherical_path(4.0, 40.0, 8, None)
    np.savez_compressed(os.path.join(OUT, "ref_rays.npz"), **out)


def gen_sampling_composite(ru, tf):
    rng = np.random.default_rng(20261017)
    v = scene.synthetic_view(6, 8, view=2)
    B = 48
    out = {"rays_o": v["rays_o"], "rays_d": v["rays_d"], "near": v["near"], "far": v["far"]}
    for lin_inv in (True, False):
        tag = "inv" if lin_inv else "lin"
        params = tf_shim.Params(perturb=True, lin_inv_depth=lin_inv)
        uc = rng.random((B, 64), dtype=F32)
        uf = rng.random((B, 128), dtype=F32)
        uf[0, :4] = [0.0, 1.0 - 2 ** -24, 0.5, 0.25]          # u = 0 and u just below 1
        tf.random.queue = [uc, uf]
        d = ru.create_input_batch_coarse_model(params, v["rays_o"], v["rays_d"], v["near"], v["far"])
        # synthetic network outputs: mix of empty space, thin surfaces and dense fog
        sig = rng.random((B * 64, 1), dtype=F32) * F32(40.0)
        sig = (sig * (rng.random((B * 64, 1)) > 0.7)).astype(F32)
        sig.reshape(B, 64)[1] = 0.0                            # fully empty ray
        sig.reshape(B, 64)[2] = 1e4                            # saturates at first sample
        sig.reshape(B, 64)[3, :] = 0.0
        sig.reshape(B, 64)[3, 17] = 500.0                      # one-hot weights
        sig.reshape(B, 64)[4, -1] = 1e-9                       # sigma_last epsilon
        rgb = rng.random((B * 64, 3), dtype=F32)
        out.update({f"{tag}_u_coarse": uc, f"{tag}_u_fine": uf, f"{tag}_t_coarse": d["t_vals"],
                    f"{tag}_bin_edges": d["bin_data"]["bin_edges"],
                    f"{tag}_xyz_coarse": d["xyz_inputs"], f"{tag}_sigma": sig, f"{tag}_rgb": rgb})
        for wb in (True, False):
            pp = ru.post_process_model_output(rgb, sig, d["t_vals"], wb)
            for k, a in pp.items():
                out[f"{tag}_wb{int(wb)}_{k}"] = a
        f = ru.create_input_batch_fine_model(params, v["rays_o"], v["rays_d"], pp["weights"],
                                             d["bin_data"], d["t_vals"])
        out[f"{tag}_t_fine_sorted"] = f["t_vals"]
        out[f"{tag}_xyz_fine"] = f["xyz_inputs"]
        dbg = rm.create_input_batch_fine_model(v["rays_o"], v["rays_d"], pp["weights"],
                                               d["bin_data"], d["t_vals"], uf, return_debug=True)
        assert np.array_equal(dbg["t_vals"], f["t_vals"])
        out[f"{tag}_cdf"] = dbg["cdf"]
        out[f"{tag}_piece_idxs"] = dbg["piece_idxs"]
        out[f"{tag}_t_fine_unsorted"] = dbg["t_vals_fine"]
        # fine-shaped compositing (S = 192) on the sorted t
        sig_f = (rng.random((B * 192, 1), dtype=F32) * F32(25.0) * (rng.random((B * 192, 1)) > 0.6)).astype(F32)
        rgb_f = rng.random((B * 192, 3), dtype=F32)
        out[f"{tag}_sigma_f"] = sig_f
        out[f"{tag}_rgb_f"] = rgb_f
        ppf = ru.post_process_model_output(rgb_f, sig_f, f["t_vals"], True)
        for k, a in ppf.items():
            out[f"{tag}_fine_wb1_{k}"] = a
    np.savez_compressed(os.path.join(OUT, "ref_sampling_composite.npz"), **out)


def gen_mlp():
    rng = np.random.default_rng(5)
    w = model.init_weights(11, bias_scale=0.05)
    R = 384
    xyz = rng.uniform(-1, 1, size=(R, 3)).astype(F32)
    d = rng.normal(size=(R, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(F32)
    import torch
    out = {"xyz": xyz, "dirs": d, "weights_seed": 11, "bias_scale": 0.05}
    for m in ("coarse", "fine"):
        r32, s32 = model.mlp_forward_np(w, m, xyz, d, torch.float32)
        r64, s64 = model.mlp_forward_np(w, m, xyz, d, torch.float64)
        out.update({f"{m}_rgb_f32": r32, f"{m}_sigma_f32": s32, f"{m}_rgb_f64": r64,
                    f"{m}_sigma_f64": s64})
    enc = model.positional_encode(torch.from_numpy(xyz), 10).numpy()
    out["enc_xyz_L10"] = enc
    out["enc_dir_L4"] = model.positional_encode(torch.from_numpy(d), 4).numpy()
    np.savez_compressed(os.path.join(OUT, "oracle_mlp.npz"), **out)


def gen_forward_and_train():
    import torch
    rng = np.random.default_rng(99)
    v = scene.synthetic_view(8, 8, view=3)
    B = 64
    uf = rng.random((B, 128), dtype=F32)
    out = {"rays_o": v["rays_o"], "rays_d": v["rays_d"], "near": v["near"], "far": v["far"],
           "u_fine": uf}
    for tag, gain in (("g1", 1.0), ("g300", 300.0)):
        w = model.init_weights(7, sigma_gain=gain)
        pc, pf, dbg = model.forward(w, v["rays_o"], v["rays_d"], v["near"], v["far"], u_fine=uf,
                                    perturb=False, white_bg=True, return_debug=True)
        pc64, pf64 = model.forward(w, v["rays_o"], v["rays_d"], v["near"], v["far"], u_fine=uf,
                                   perturb=False, white_bg=True, mlp_dtype=torch.float64)
        for k in ("pred_rgb", "pred_depth", "acc_map"):
            out[f"{tag}_c_{k}"] = pc[k]
            out[f"{tag}_f_{k}"] = pf[k]
            out[f"{tag}_f64_{k}"] = pf64[k]
        out[f"{tag}_c_weights"] = pc["weights"]
        out[f"{tag}_t_fine_sorted"] = dbg["t_fine_sorted"]
    # one training step (perturb off, fixed uniforms), fp32 and fp64 autograd
    gt = rng.random((B, 3), dtype=F32)
    out["rgb_gt"] = gt
    w = model.init_weights(7)
    names = model.all_variable_names()
    m = {n: np.zeros_like(w[n]) for n in names}
    vv = {n: np.zeros_like(w[n]) for n in names}
    batch = ((v["rays_o"], v["rays_d"], v["near"], v["far"]), (gt,))
    it, info = model.train_step(w, m, vv, 0, batch, u_fine=uf, perturb=False, white_bg=True)
    out["train_loss"] = info["loss"]
    out["train_coarse_loss"] = info["coarse_loss"]
    out["train_fine_loss"] = info["fine_loss"]
    out["train_psnr_metric"] = info["psnr_metric"]
    out["train_grad_norms"] = np.array([np.linalg.norm(info["grads"][n].astype(np.float64)) for n in names])
    out["train_grad_fine_dense_9_bias"] = info["grads"]["fine/dense_9/bias"]
    out["train_grad_coarse_rgb_kernel"] = info["grads"]["coarse/rgb/kernel"]
    out["train_grad_fine_sigma_kernel"] = info["grads"]["fine/sigma/kernel"]
    out["train_param_after_fine_rgb_kernel"] = w["fine/rgb/kernel"]
    out["train_param_after_coarse_dense_0_bias"] = w["coarse/dense_0/bias"]
    w64 = model.init_weights(7)
    info64, g64 = model.loss_and_grads(w64, v["rays_o"], v["rays_d"], v["near"], v["far"], gt,
                                       u_fine=uf, perturb=False, white_bg=True, dtype=torch.float64)
    out["train_loss_f64"] = info64["loss"]
    out["train_grad_norms_f64"] = np.array([np.linalg.norm(g64[n].astype(np.float64)) for n in names])
    np.savez_compressed(os.path.join(OUT, "oracle_forward_train.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    with tf_shim.reference_ray_utils() as (ru, pu, tf):
        gen_rays(ru, pu)
        gen_sampling_composite(ru, tf)
    gen_mlp()
    gen_forward_and_train()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
