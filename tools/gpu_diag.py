"""Developer diagnostics for the tensor-core MLP kernel: run on the GPU box, prints error
statistics against the on-device fp32 path (not a test, not a benchmark)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import nerf_tf2_b200 as nb
from nerf_tf2_b200 import _lib


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    p = nb.make_params({"system": {"white_bg": True}}, perturb=False)
    nerf = nb.setup_model(p, precision="bf16", seed=3)
    rng = np.random.default_rng(0)
    # small random biases so that bias handling is exercised
    flat = nerf.flat_params.cpu().numpy().copy()
    for v in nerf.trainable_variables:
        if v.name.endswith("bias"):
            flat[v._ofs:v._ofs + v._n] = rng.uniform(-0.05, 0.05, v._n).astype(np.float32)
    nerf.set_flat_params(flat)
    for R in (128, 256, 1000, 40000):
        xyz = torch.from_numpy(rng.uniform(-1, 1, (R, 3)).astype(np.float32)).cuda()
        d = rng.normal(size=(R, 3)); d = torch.from_numpy((d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)).cuda()
        rgb32, sig32 = nerf.coarse_model((xyz, d), precision=_lib.FP32)
        torch.cuda.synchronize()
        for name, prec in (("bf16", _lib.BF16), ("fp16", _lib.FP16)):
            t0 = time.time()
            rgb, sig = nerf.coarse_model((xyz, d), precision=prec)
            torch.cuda.synchronize()
            dt = time.time() - t0
            e = (rgb - rgb32).abs()
            es = (sig - sig32).abs()
            print(f"R={R:6d} {name}: rgb err p50 {e.median():.2e} p99 {e.flatten().kthvalue(max(1,int(0.99*e.numel())))[0]:.2e} "
                  f"max {e.max():.2e} | sigma err max {es.max():.2e} (sigma max {sig32.max():.3f}) nan={int(torch.isnan(rgb).sum())} "
                  f"t={dt*1e3:.2f} ms", flush=True)
            if R == 128 and float(e.max()) > 0.1:
                print("rgb tc  :", rgb[:4].cpu().numpy())
                print("rgb fp32:", rgb32[:4].cpu().numpy())
                print("sig tc  :", sig[:8, 0].cpu().numpy())
                print("sig fp32:", sig32[:8, 0].cpu().numpy())
    # throughput probe on a render-sized problem
    B, S = 65536, 192
    ro = torch.zeros((B, 3), device="cuda"); rd = torch.nn.functional.normalize(torch.randn((B, 3), device="cuda"), dim=1)
    t = torch.sort(torch.rand((B, S), device="cuda") * 0.85 + 0.425, dim=1)[0].contiguous()
    for name, prec in (("bf16", _lib.BF16), ("fp16", _lib.FP16)):
        for _ in range(2):
            nerf._mlp(1, ro, rd, t, prec)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3):
            nerf._mlp(1, ro, rd, t, prec)
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 3
        flop = B * S * 1186816
        print(f"fused MLP {name}: {B*S} rows in {ms:.3f} ms -> {flop/ms/1e9:.1f} TFLOP/s algorithmic", flush=True)


if __name__ == "__main__":
    main()
