"""HBM-bound kernels timed alone: the integrator (coarse 64 samples with the weights output, fine 192 samples
without/with it) and the hierarchical sampler (in-kernel Philox and explicit uniforms) at the render chunk
size, back to back over rotating input sets larger than L2. Prints one JSON object per kernel:
algorithmic bytes per ray (SURVEY.md section 8d / DESIGN.md section 4), GB/s, fraction of the measured HBM peak.

    python tools/hbm_kernels_bench.py [--rays 65536] [--iters 40] [--out gpurun_out/hbm_kernels.json]
    ncu --set full -k regex:composite_fwd\\|sample_fine -c 6 python tools/hbm_kernels_bench.py --iters 1 --warmup 0 --sets 1
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_tf2_b200 as nb  # noqa: E402
from nerf_tf2_b200 import ray_utils as ru  # noqa: E402
from nerf_tf2_b200._lib import load, ptr, stream_ptr, check  # noqa: E402


def hbm_peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6500.0, "fallback"


def timed(fn, sets, iters, warmup):
    for i in range(warmup):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--sets", type=int, default=0, help="input sets to rotate through (0: enough to exceed 2x L2)")
    ap.add_argument("--nc", type=int, default=64)
    ap.add_argument("--nf", type=int, default=128)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    B, Nc, Nf = a.rays, a.nc, a.nf
    S = Nc + Nf
    peak, src = hbm_peak()
    g = torch.Generator(device="cuda").manual_seed(0)
    L2 = 126e6
    results = []

    def nsets(bytes_per_set):
        return a.sets if a.sets else max(2, int(2 * L2 / bytes_per_set) + 1)

    def comp_sets(s, n):
        out = []
        for _ in range(n):
            t = torch.sort(torch.rand((B, s), device="cuda", generator=g) * 0.8 + 0.4, dim=1).values.contiguous()
            sig = torch.rand((B * s,), device="cuda", generator=g) * 20 * (torch.rand((B * s,), device="cuda", generator=g) > 0.6)
            rgb = torch.rand((B * s, 3), device="cuda", generator=g)
            out.append((rgb, sig, t))
        return out

    for name, s, need_w in (("composite_fwd coarse (weights out)", Nc, True),
                            ("composite_fwd fine (render: no weights)", S, False),
                            ("composite_fwd fine (weights out)", S, True)):
        bpr = (24 if need_w else 20) * s + 20
        sets = comp_sets(s, nsets(B * bpr))
        # outputs preallocated, C ABI called directly: a ~25 us kernel is otherwise bounded by the Python-side
        # allocations of ray_utils.post_process_model_output, not by the GPU
        wts = torch.empty((B, s), device="cuda") if need_w else None
        prgb = torch.empty((B, 3), device="cuda"); pdep = torch.empty((B,), device="cuda"); pacc = torch.empty((B,), device="cuda")
        lib, st = load(), stream_ptr()

        def run_comp(x):
            check(lib.nerfb200_composite_fwd(B, s, ptr(x[1]), ptr(x[0]), ptr(x[2]), 1, ptr(wts, allow_none=True), ptr(prgb),
                                             ptr(pdep), ptr(pacc), st), "composite_fwd")
        ms = timed(run_comp, sets, a.iters, a.warmup)
        gbs = B * bpr / (ms / 1e3) / 1e9
        results.append({"kernel": name, "rays": B, "S": s, "bytes_per_ray": bpr, "us": ms * 1e3, "GBps": gbs,
                        "frac_of_hbm_peak": gbs / peak, "input_sets": len(sets)})
        del sets

    # sampler inputs: weights from a real integrator pass over random sigma (bimodal, like a random-init network)
    def samp_sets(n, with_u):
        out = []
        near = torch.full((B,), 0.425, device="cuda"); far = torch.full((B,), 1.275, device="cuda")
        for i in range(n):
            t_c, edges = ru.sample_coarse(Nc, True, True, near, far, None, seed=i, ray0=0)
            sig = torch.rand((B * Nc,), device="cuda", generator=g) * 20 * (torch.rand((B * Nc,), device="cuda", generator=g) > 0.6)
            w = ru.compute_weights(sig, t_c)
            u = torch.rand((B, Nf), device="cuda", generator=g) if with_u else None
            out.append((w, edges, t_c, u))
        return out

    for name, with_u in (("sample_fine (in-kernel Philox)", False), ("sample_fine (explicit u)", True)):
        # weights + bin edges + t_coarse read, t_sorted written (+ u read)
        bpr = 4 * Nc * 2 + 4 * (Nc + 1) + 4 * S + (4 * Nf if with_u else 0)
        sets = samp_sets(nsets(B * bpr), with_u)
        tso = torch.empty((B, S), device="cuda")
        lib, st = load(), stream_ptr()

        def run_samp(x):
            check(lib.nerfb200_sample_fine(B, Nc, Nf, ptr(x[0]), ptr(x[1]), ptr(x[2]), ptr(x[3], allow_none=True), 1, 0,
                                           ptr(tso), None, None, None, st), "sample_fine")
        ms = timed(run_samp, sets, a.iters, a.warmup)
        gbs = B * bpr / (ms / 1e3) / 1e9
        results.append({"kernel": name, "rays": B, "Nc": Nc, "Nf": Nf, "bytes_per_ray": bpr, "us": ms * 1e3, "GBps": gbs,
                        "frac_of_hbm_peak": gbs / peak, "input_sets": len(sets)})
        del sets

    for r in results:
        r["hbm_peak_GBps"] = peak
        r["peak_source"] = src
        print(json.dumps(r))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
