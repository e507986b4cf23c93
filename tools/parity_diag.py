"""Developer tool (GPU box): error statistics of the tensor-core render against the fp32 oracle on a full cfg1 view
(BASELINE.json configs[0]: 100x100, 64+128 samples), per precision, with and without the split last-sample launch,
plus 800x800 render timings. Output feeds the tolerances stated in tests/ and DESIGN.md section 2."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import nerf_tf2_b200 as nb
from oracle import model as om, scene as osc

H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 100
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/parity_diag.json"
F32 = np.float32
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
res = {"view": f"{H}x{W}", "cases": []}
for gain in (1.0, 300.0):
    v = osc.synthetic_view(H, W, view=1)
    rng = np.random.default_rng(11)
    uf = rng.random((H * W, 128), dtype=F32)
    w = om.init_weights(7, sigma_gain=gain)
    t0 = time.time()
    pc, pf, dbg = om.forward(w, v["rays_o"], v["rays_d"], v["near"], v["far"], u_fine=uf, perturb=False, white_bg=True, return_debug=True)
    res["oracle_s"] = time.time() - t0
    last_c = dbg["sigma_c"].reshape(H * W, 64)[:, -1] > 0
    last_f = dbg["sigma_f"].reshape(H * W, 192)[:, -1] > 0
    for prec in ("bf16", "fp16", "tf32", "fp32"):
        for precise in ((True, False) if prec != "fp32" else (True,)):
            nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=False), precision=prec, precise_last=precise)
            nerf.set_weights_from_dict(w)
            tr = {}
            oc, of = nerf.forward(dev(v["rays_o"]), dev(v["rays_d"]), dev(v["near"]), dev(v["far"]), u_fine=dev(uf), _train=None)
            torch.cuda.synchronize()
            st = {"precision": prec, "precise_last": precise, "sigma_gain": gain}
            for name, o, r in (("coarse", oc, pc), ("fine", of, pf)):
                for k in ("pred_rgb", "pred_depth", "acc_map"):
                    e = np.abs(o[k].cpu().numpy().reshape(H * W, -1) - r[k].reshape(H * W, -1)).max(axis=1)
                    st[f"{name}_{k}"] = {"p50": float(np.percentile(e, 50)), "p99": float(np.percentile(e, 99)),
                                         "p999": float(np.percentile(e, 99.9)), "max": float(e.max()),
                                         "frac_gt_3e-2": float((e > 3e-2).mean())}
            res["cases"].append(st)
            print(json.dumps(st), flush=True)
            del nerf
# render timing at 800x800 (device-resident rays)
v8 = osc.synthetic_view(800, 800, view=0)
args = [dev(v8[k]) for k in ("rays_o", "rays_d", "near", "far")]
res["timing"] = []
for prec, precise in (("bf16", True), ("bf16", False), ("fp16", True), ("tf32", True), ("tf32", False)):
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}), precision=prec, precise_last=precise, render_chunk=65536)
    for _ in range(2):
        nerf.render_rays(*args, need_weights=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 4
    for _ in range(n):
        nerf.render_rays(*args, need_weights=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t = {"precision": prec, "precise_last": precise, "ms_per_view": ms, "rays_per_s": 640000 / ms * 1e3,
         "mlp_tflops": 640000 * 256 * 1186816 / (ms * 1e-3) / 1e12}
    res["timing"].append(t)
    print(json.dumps(t), flush=True)
    del nerf
json.dump(res, open(out_path, "w"), indent=1)
