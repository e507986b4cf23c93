"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for `nerf/core/model.py`.

torch-CPU restatement (fp32 by default, fp64 on request to separate "whose error is
whose") of the reference's PositionalEncoder (core/model.py:289-332), the 8x256 MLP
(core/model.py:334-394), NeRF.forward (:57-125), train_step (:127-180) and the
Adam + ExponentialDecay optimiser set up in setup_model (:413-418).

PARITY UNPINNED for this file: the arithmetic lives in tensorflow==2.6.0 /
keras==2.7.0 (requirements.txt:44,17), which cannot be installed here, and the
reference ships no golden vectors. Keras Dense = act(x @ kernel[in,out] + bias);
Keras Adam (OptimizerV2, non-amsgrad, epsilon=1e-7) as in SURVEY.md Appendix A.
Cross-checked (not pinned) against independent implementations of the same published semantics:
tests/test_oracle_crosscheck.py (scikit-learn's Adam and MLP forward, torch linear, fp64 finite differences).
"""
import math

import numpy as np
import torch

from . import ray_march as rm

F32 = np.float32

LAYER_NAMES = [f"dense_{i}" for i in range(10)] + ["rgb", "sigma"]
# Order of Keras' `model.layers` (hence trainable_variables / get_weights / optimizer slots) for the
# functional model of core/model.py:334-394: Functional._map_graph_network sorts layers by decreasing
# depth from the outputs [rgb, sigma] and breaks ties by the output-first traversal index, so the two
# heads come last, rgb before sigma. (TensorFlow is not installable here: restated, not executed;
# tests/test_keras_layer_order.py runs the restated algorithm on the reference's graph.)
# (fan_in, fan_out) per layer (core/model.py:366-387)
LAYER_SHAPES = {
    "dense_0": (63, 256), "dense_1": (256, 256), "dense_2": (256, 256), "dense_3": (256, 256),
    "dense_4": (256, 256), "dense_5": (319, 256), "dense_6": (256, 256), "dense_7": (256, 256),
    "sigma": (256, 1), "dense_8": (256, 256), "dense_9": (283, 128), "rgb": (128, 3),
}


def variable_names(model_name):
    out = []
    for ln in LAYER_NAMES:
        out += [f"{model_name}/{ln}/kernel", f"{model_name}/{ln}/bias"]
    return out


def all_variable_names():
    """nerf.trainable_variables order: coarse then fine, each in LAYER_NAMES order."""
    return variable_names("coarse") + variable_names("fine")


def init_weights(seed, bias_scale=0.0, sigma_gain=1.0):
    """Keras defaults: glorot_uniform kernels, zero biases. `bias_scale`>0 draws small
    random biases instead (tests only, so that bias handling is exercised);
    `sigma_gain` scales the sigma kernel (SURVEY.md 8d 'sharpened' variant)."""
    rng = np.random.default_rng(seed)
    w = {}
    for m in ("coarse", "fine"):
        for ln in LAYER_NAMES:
            fi, fo = LAYER_SHAPES[ln]
            lim = math.sqrt(6.0 / (fi + fo))
            k = rng.uniform(-lim, lim, size=(fi, fo)).astype(F32)
            if ln == "sigma":
                k = (k * F32(sigma_gain)).astype(F32)
            w[f"{m}/{ln}/kernel"] = k
            if bias_scale > 0:
                w[f"{m}/{ln}/bias"] = rng.uniform(-bias_scale, bias_scale, size=(fo,)).astype(F32)
            else:
                w[f"{m}/{ln}/bias"] = np.zeros((fo,), dtype=F32)
    return w


def positional_encode(x, L):
    """PositionalEncoder.call (core/model.py:305-332). x: torch [R,3]."""
    mult = (np.float32(2.0) ** np.arange(L, dtype=F32)) * F32(np.pi)   # fl32(2^l)*fl32(pi)
    mult_t = torch.from_numpy(mult.astype(F32)).to(x.dtype)             # exact in fp64 too
    expanded = x[..., None] * mult_t.reshape(1, 1, -1)
    inter = torch.stack([torch.sin(expanded), torch.cos(expanded)], dim=-1)
    sincos = inter.reshape(-1, x.shape[1] * L * 2)
    return torch.cat([x, sincos], dim=-1)


def mlp_forward(w, model_name, xyz, dirs, return_pre=False):
    """get_coarse_or_fine_model forward (core/model.py:334-394). w: name -> torch tensor. `return_pre` also returns
    the sigma head's pre-activation (the reference applies ReLU to it, core/model.py:375): the distance of the LAST
    sample's pre-activation from zero says how well conditioned a ray's output is (delta_last = 1e10,
    utils/ray_utils.py:459-468: alpha_last is 0 or 1 according to its sign)."""
    def dense(name, h):
        return h @ w[f"{model_name}/{name}/kernel"] + w[f"{model_name}/{name}/bias"]

    enc_xyz = positional_encode(xyz, 10)
    enc_dir = positional_encode(dirs, 4)
    h = enc_xyz
    for i in range(8):
        h = torch.relu(dense(f"dense_{i}", h))
        if i == 4:
            h = torch.cat([h, enc_xyz], dim=-1)
    sigma_pre = dense("sigma", h)
    sigma = torch.relu(sigma_pre)
    bott = dense("dense_8", h)
    g = torch.cat([bott, enc_dir], dim=-1)
    g = torch.relu(dense("dense_9", g))
    rgb = torch.sigmoid(dense("rgb", g))
    if return_pre:
        return rgb, sigma, sigma_pre
    return rgb, sigma


def to_torch(w, dtype=torch.float32, requires_grad=False):
    out = {}
    for k, v in w.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
        if requires_grad:
            t.requires_grad_(True)
        out[k] = t
    return out


def mlp_forward_np(w_np, model_name, xyz, dirs, dtype=torch.float32, chunk=1 << 16, return_pre=False):
    """NumPy-in / NumPy-out chunked forward (no grad)."""
    w = to_torch(w_np, dtype)
    rgbs, sigmas, pres = [], [], []
    with torch.no_grad():
        for s in range(0, xyz.shape[0], chunk):
            r, sg, pre = mlp_forward(w, model_name,
                                     torch.from_numpy(np.ascontiguousarray(xyz[s:s + chunk])).to(dtype),
                                     torch.from_numpy(np.ascontiguousarray(dirs[s:s + chunk])).to(dtype), return_pre=True)
            rgbs.append(r.to(torch.float32).numpy())
            sigmas.append(sg.to(torch.float32).numpy())
            pres.append(pre.to(torch.float32).numpy())
    if return_pre:
        return np.concatenate(rgbs, 0), np.concatenate(sigmas, 0), np.concatenate(pres, 0)
    return np.concatenate(rgbs, 0), np.concatenate(sigmas, 0)


def forward(w_np, rays_o, rays_d, near, far, N_coarse=64, N_fine=128, lin_inv_depth=True,
            perturb=False, white_bg=True, u_coarse=None, u_fine=None, mlp_dtype=torch.float32,
            return_debug=False):
    """NeRF.forward (core/model.py:57-125) with explicit uniforms."""
    d_cm = rm.create_input_batch_coarse_model(N_coarse, lin_inv_depth, perturb, rays_o, rays_d,
                                              near, far, u_coarse)
    rgb_c, sig_c, pre_c = mlp_forward_np(w_np, "coarse", d_cm["xyz_inputs"], d_cm["dir_inputs"], mlp_dtype, return_pre=True)
    pp_c = rm.post_process_model_output(rgb_c, sig_c, d_cm["t_vals"], white_bg)
    d_fm = rm.create_input_batch_fine_model(rays_o, rays_d, pp_c["weights"], d_cm["bin_data"],
                                            d_cm["t_vals"], u_fine, return_debug=True)
    rgb_f, sig_f, pre_f = mlp_forward_np(w_np, "fine", d_fm["xyz_inputs"], d_fm["dir_inputs"], mlp_dtype, return_pre=True)
    pp_f = rm.post_process_model_output(rgb_f, sig_f, d_fm["t_vals"], white_bg)
    if return_debug:
        dbg = {"t_coarse": d_cm["t_vals"], "t_fine_sorted": d_fm["t_vals"],
               "rgb_c": rgb_c, "sigma_c": sig_c, "rgb_f": rgb_f, "sigma_f": sig_f,
               "cdf": d_fm["cdf"], "piece_idxs": d_fm["piece_idxs"],
               "bin_edges": d_cm["bin_data"]["bin_edges"],
               # pre-activation of the sigma head at the LAST sample of every ray, coarse and fine
               "sigma_pre_last_c": pre_c.reshape(d_cm["t_vals"].shape)[:, -1].copy(),
               "sigma_pre_last_f": pre_f.reshape(d_fm["t_vals"].shape)[:, -1].copy()}
        return pp_c, pp_f, dbg
    return pp_c, pp_f


# ---------------------------------------------------------------- training (autograd)
def composite_torch(rgb, sigma, t_vals, white_bg):
    """post_process_model_output (utils/ray_utils.py:484-551) in torch for autograd."""
    B, S = t_vals.shape
    diffs = t_vals[:, 1:] - t_vals[:, :-1]
    diffs = torch.cat([diffs, torch.full((B, 1), 1e10, dtype=t_vals.dtype)], dim=-1)
    sig = sigma.reshape(B, S)
    alpha = 1 - torch.exp(-sig * diffs)
    trans = torch.cumprod(1 - alpha + 1e-10, dim=1)
    trans = torch.cat([torch.ones((B, 1), dtype=t_vals.dtype), trans[:, :-1]], dim=1)
    weights = alpha * trans
    pred_rgb = torch.sum(weights[..., None] * rgb.reshape(B, S, 3), dim=1)
    acc = torch.sum(weights, dim=1)
    depth = torch.sum(weights * t_vals, dim=1)
    if white_bg:
        pred_rgb = pred_rgb + (1 - acc[:, None])
    return {"weights": weights, "pred_rgb": pred_rgb, "acc_map": acc, "pred_depth": depth}


def loss_and_grads(w_np, rays_o, rays_d, near, far, rgb_gt, N_coarse=64, N_fine=128,
                   lin_inv_depth=True, perturb=False, white_bg=True, u_coarse=None, u_fine=None,
                   dtype=torch.float32):
    """Forward + tape.gradient part of NeRF.train_step (core/model.py:148-170).
    Sample positions are constants for autograd (stop_gradient, utils/ray_utils.py:377)."""
    w = to_torch(w_np, dtype, requires_grad=True)
    d_cm = rm.create_input_batch_coarse_model(N_coarse, lin_inv_depth, perturb, rays_o, rays_d,
                                              near, far, u_coarse)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    rgb_c, sig_c = mlp_forward(w, "coarse", tt(d_cm["xyz_inputs"]), tt(d_cm["dir_inputs"]))
    pp_c = composite_torch(rgb_c, sig_c, tt(d_cm["t_vals"]), white_bg)
    w_c_np = pp_c["weights"].detach().to(torch.float32).numpy()
    d_fm = rm.create_input_batch_fine_model(rays_o, rays_d, w_c_np, d_cm["bin_data"],
                                            d_cm["t_vals"], u_fine)
    rgb_f, sig_f = mlp_forward(w, "fine", tt(d_fm["xyz_inputs"]), tt(d_fm["dir_inputs"]))
    pp_f = composite_torch(rgb_f, sig_f, tt(d_fm["t_vals"]), white_bg)
    gt = tt(rgb_gt)
    coarse_loss = torch.mean((gt - pp_c["pred_rgb"]) ** 2)       # Keras MeanSquaredError
    fine_loss = torch.mean((gt - pp_f["pred_rgb"]) ** 2)
    total = coarse_loss + fine_loss
    names = all_variable_names()
    grads = torch.autograd.grad(total, [w[n] for n in names])
    g = {n: gr.detach().to(torch.float32).numpy() for n, gr in zip(names, grads)}
    out = {"loss": float(total.detach()), "coarse_loss": float(coarse_loss.detach()),
           "fine_loss": float(fine_loss.detach()),
           "pred_rgb_c": pp_c["pred_rgb"].detach().to(torch.float32).numpy(),
           "pred_rgb_f": pp_f["pred_rgb"].detach().to(torch.float32).numpy(),
           "t_fine_sorted": d_fm["t_vals"], "weights_c": w_c_np}
    return out, g


def exponential_decay_lr(step, initial=5e-4, decay_steps=500000, decay_rate=0.1):
    """ExponentialDecay (core/model.py:413-417), staircase=False."""
    return initial * (decay_rate ** (step / decay_steps))


def adam_step(w_np, grads, m, v, iterations, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras OptimizerV2 Adam._resource_apply_dense, non-amsgrad.
    `iterations` is the value BEFORE this step. fp32 arithmetic. In-place on w_np/m/v."""
    t = iterations + 1
    lr = F32(exponential_decay_lr(iterations))
    lr_t = F32(lr * F32(math.sqrt(1.0 - beta2 ** t)) / F32(1.0 - beta1 ** t))
    b1, b2 = F32(beta1), F32(beta2)
    for n in grads:
        g = grads[n].astype(F32)
        m[n] = (m[n] * b1 + g * (F32(1) - b1)).astype(F32)
        v[n] = (v[n] * b2 + (g * g) * (F32(1) - b2)).astype(F32)
        w_np[n] = (w_np[n] - lr_t * m[n] / (np.sqrt(v[n]) + F32(eps))).astype(F32)
    return iterations + 1


def train_step(w_np, m, v, iterations, batch, **kw):
    """NeRF.train_step (core/model.py:127-180): returns (new_iterations, info)."""
    (rays_o, rays_d, near, far), (rgb,) = batch
    info, g = loss_and_grads(w_np, rays_o, rays_d, near, far, rgb, **kw)
    it = adam_step(w_np, g, m, v, iterations)
    metric = rm.PSNRMetric()
    metric.update_state(rgb, info["pred_rgb_f"])
    info["psnr_metric"] = float(metric.result())
    info["grads"] = g
    return it, info
