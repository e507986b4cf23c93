"""Developer tool (GPU box): accuracy of the split last-sample launch. For a view, runs the bf16 ray march step by step
through the C ABI, then recomputes sigma of the LAST sample of every ray (coarse and fine, at the GPU's own sample
positions) in fp64 on the CPU and reports the error of the GPU value and every sign mismatch."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import nerf_tf2_b200 as nb
from nerf_tf2_b200 import ray_utils as ru, _lib
from oracle import model as om, scene as osc


def sigma_pre64(w_np, model, xyz, dirs):
    wt = om.to_torch(w_np, torch.float64)
    x = torch.from_numpy(xyz).double()
    enc = om.positional_encode(x, 10)
    h = enc
    for i in range(8):
        h = torch.relu(h @ wt[f"{model}/dense_{i}/kernel"] + wt[f"{model}/dense_{i}/bias"])
        if i == 4:
            h = torch.cat([h, enc], -1)
    return (h @ wt[f"{model}/sigma/kernel"] + wt[f"{model}/sigma/bias"]).numpy().reshape(-1)


def run(H, W, view, wseed, prec="bf16"):
    v = osc.synthetic_view(H, W, view=view)
    w = om.init_weights(wseed)
    n = H * W
    uf = np.random.default_rng(0).random((n, 128), dtype=np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ro, rd, near, far = (dev(v[k]) for k in ("rays_o", "rays_d", "near", "far"))
    for precise in (True, False):
        nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=False), precision=prec, precise_last=precise)
        nerf.set_weights_from_dict(w)
        t_c, edges = ru.sample_coarse(64, True, False, near, far, None, 0, 0)
        rgb_c, sig_c = nerf._mlp(0, ro, rd, t_c)
        pp_c = ru.post_process_model_output(rgb_c, sig_c, t_c, True)
        t_f = ru.sample_fine(128, pp_c["weights"], edges, t_c, dev(uf), 0, 0)
        rgb_f, sig_f = nerf._mlp(1, ro, rd, t_f)
        torch.cuda.synchronize()
        for model, t, sig, S in (("coarse", t_c, sig_c, 64), ("fine", t_f, sig_f, 192)):
            tl = t[:, -1:].cpu().numpy()
            xyz = (v["rays_o"] + tl * v["rays_d"]).astype(np.float32)
            ref = sigma_pre64(w, model, xyz, v["rays_d"])
            got = sig.reshape(n, S)[:, -1].cpu().numpy()
            err = np.abs(got - np.maximum(ref, 0))
            mism = np.nonzero((got > 0) != (ref > 0))[0]
            print(f"{H}x{W} view {view} seed {wseed} {prec} precise={precise} {model}: |sigma_last - relu(fp64)| median {np.median(err):.2e} "
                  f"p99 {np.percentile(err, 99):.2e} max {err.max():.2e} (ray {err.argmax()}, tile {err.argmax() // 128}, row {err.argmax() % 128}); "
                  f"sign mismatches {len(mism)}: " + ", ".join(f"ray {i} ref {ref[i]:.2e} got {got[i]:.2e}" for i in mism[:6]), flush=True)


if __name__ == "__main__":
    run(24, 24, 2, 5)
    run(24, 24, 2, 5, "fp16")
    run(40, 40, 1, 7)
    run(100, 100, 1, 7)
