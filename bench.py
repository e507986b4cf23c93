#!/usr/bin/env python
"""
bench.py -- headline benchmark of the NeRF ray-march hot path (BASELINE.json `metric`:
rendered rays/s, 64+128 samples, 800x800).

  python bench.py --gpus N --steps K --warmup W            (this repo's sm_100a path)
  python bench.py --impl reference --gpus N --steps K ...   (the reference's CPU path: the oracle
                                                             port, TensorFlow is not installable)

A "step" is one full 800x800 view (640 000 rays, 157 reference chunks' worth of work) marched
through coarse + fine networks. With N > 1 (torchrun, one rank per GPU) every rank renders its own
view of the synthetic 360-degree scene -- views are independent units, no data-path collective --
so scaling is "weak" and `value` is the whole-job aggregate.

One JSON line is printed by rank 0. Besides the contract keys it carries
  roofline      dominant kernel (fused encoding+MLP, tensor bound): algorithmic FLOPs / CUDA-event time
  roofline_hbm  the integrator and the hierarchical sampler against measured HBM bandwidth
  cpu_baseline  the oracle timed on the host cores on a bounded sample (rank 0, N=1)
  e2e           the same metric through NeRF.predict() with HOST rays (H2D + D2H inside the timing)
  train         4096-ray coarse+fine train steps/s (secondary metric of BASELINE.json)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 800
N_COARSE, N_FINE = 64, 128
FLOP_PER_ROW = 1186816                      # 2 x 593 408 MAC, unpadded
ROWS_PER_RAY = N_COARSE + (N_COARSE + N_FINE)
# `ncu -i <report> --page raw --csv` exports of this round's `ncu --set full` captures (traffic = dram bytes per launch)
NCU_MLP_CSV = {"bf16": "r2_ncu_full_mlp_pair_raw.csv", "fp16": "r2_ncu_full_mlp_pair_raw.csv", "tf32": "r2_ncu_full_mlp_tf32_raw.csv"}
NCU_HBM_CSV = "r2_ncu_full_hbm_kernels_raw.csv"
REF_SAMPLE_RAYS = 1024                      # bounded sample per step for the CPU arm
METRIC = "rendered rays/s (64+128 samples, 800x800)"
WORKLOAD = ("BASELINE.json configs[1]: synthetic lego-shaped 800x800 single-view render, 64+128 samples, random-init "
            "coarse+fine 8x256 MLPs, white_bg, perturb on (in-kernel Philox)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


class OracleArm:
    """The reference arithmetic on the host cores: oracle forward (fixed uniforms). Scene, weights and the ray sample
    are built ONCE; a timed call is the forward only (coarse sampling -> MLP -> integrator -> hierarchical sampling ->
    MLP -> integrator), what the reference's predict_step spends its time in."""

    def __init__(self, view_hw, sample_rays=None, seed=0):
        import torch
        from oracle import model as om, scene as osc
        self.om = om
        self.threads = os.cpu_count()
        torch.set_num_threads(self.threads)
        h, w = view_hw
        v = osc.synthetic_view(h, w, view=0)
        rng = np.random.default_rng(seed)
        n = h * w
        sel = np.arange(n) if sample_rays is None or sample_rays >= n else rng.choice(n, size=sample_rays, replace=False)
        self.n = len(sel)
        self.rays = tuple(np.ascontiguousarray(v[k][sel]) for k in ("rays_o", "rays_d", "near", "far"))
        self.w = om.init_weights(0)
        self.uc = rng.random((self.n, N_COARSE), dtype=np.float32)
        self.uf = rng.random((self.n, N_FINE), dtype=np.float32)

    def step(self):
        t0 = time.time()
        self.om.forward(self.w, *self.rays, N_COARSE, N_FINE, lin_inv_depth=True, perturb=True, white_bg=True,
                        u_coarse=self.uc, u_fine=self.uf)
        return time.time() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = OracleArm((H, W), REF_SAMPLE_RAYS)
    for _ in range(args.warmup):
        arm.step()
    dt = sum(arm.step() for _ in range(args.steps))
    val = arm.n * args.steps / dt
    sample = (f"{arm.n} random rays of the 800x800 view per step (scene, weights and sample built once outside the timed "
              f"steps), coarse+fine 64+128, fp32 torch-CPU/NumPy oracle")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + " (bounded sample)"},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": arm.threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class KernelTimer:
    """CUDA-event brackets around the C-ABI launches, on the launching (current) stream."""

    def __init__(self, torch):
        self.torch, self.ev = torch, {}

    def wrap(self, name, fn):
        def inner(*a, **k):
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            self.ev.setdefault(name, []).append((e0, e1))
            return out
        return inner

    def totals(self):
        return {n: (sum(e0.elapsed_time(e1) for e0, e1 in evs), len(evs)) for n, evs in self.ev.items()}


def hbm_kernels_alone(rays=65536, iters=20, warmup=5, sets=0, nc=N_COARSE, nf=N_FINE):
    """The HBM-bound kernels of the path timed ALONE at the render chunk size: back-to-back launches through the C ABI
    with preallocated outputs, CUDA events around the batch, inputs rotated through sets that together exceed 2x L2.
    (Event brackets around single launches inside a render step add several microseconds of launch gap to kernels
    that run for 15-50 us, so the in-step figures understate them.) Returns one dict per kernel: algorithmic bytes
    per ray, microseconds per launch, GB/s."""
    import torch
    from nerf_tf2_b200 import ray_utils as ru
    from nerf_tf2_b200._lib import load, ptr, stream_ptr, check

    B, Nc, Nf, S = rays, nc, nf, nc + nf
    g = torch.Generator(device="cuda").manual_seed(0)
    L2 = 126e6
    lib, st = load(), stream_ptr()
    results = []

    def nsets(bytes_per_set):
        return sets if sets else max(2, int(2 * L2 / bytes_per_set) + 1)

    def timed(fn, data):
        for i in range(warmup):
            fn(data[i % len(data)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(data[i % len(data)])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def rnd(*shape):
        return torch.rand(shape, device="cuda", generator=g)

    def sigma_like(n):       # sigma == 0 for most samples, like a random-init network
        return rnd(n) * 20 * (rnd(n) > 0.6)

    near, far = torch.full((B,), 0.425, device="cuda"), torch.full((B,), 1.275, device="cuda")

    def sorted_t(s, seed):   # ascending sample positions from the path's own samplers (no torch.sort in the launch list)
        t_c, edges = ru.sample_coarse(Nc, True, True, near, far, None, seed=100 + seed, ray0=0)
        if s == Nc:
            return t_c
        return ru.sample_fine(Nf, ru.compute_weights(sigma_like(B * Nc), t_c), edges, t_c, None, seed=100 + seed, ray0=0)

    coarse_name = "composite_fwd_kernel<4,full,2 rays/warp>" if Nc == 64 else f"composite_fwd_kernel<{Nc // 32},full>"
    for key, name, s, need_w in (("composite_coarse", coarse_name + " (coarse, weights out)", Nc, True),
                                 ("composite_fine", f"composite_fwd_kernel<{S // 32},full> (fine, render: no weights)", S, False),
                                 ("composite_fine_w", f"composite_fwd_kernel<{S // 32},full> (fine, weights out)", S, True)):
        bpr = (24 if need_w else 20) * s + 20
        data = [(rnd(B * s, 3), sigma_like(B * s), sorted_t(s, i)) for i in range(nsets(B * bpr))]
        wts = torch.empty((B, s), device="cuda") if need_w else None
        prgb, pdep, pacc = torch.empty((B, 3), device="cuda"), torch.empty((B,), device="cuda"), torch.empty((B,), device="cuda")

        def run(x):
            check(lib.nerfb200_composite_fwd(B, s, ptr(x[1]), ptr(x[0]), ptr(x[2]), 1, ptr(wts, allow_none=True), ptr(prgb),
                                             ptr(pdep), ptr(pacc), st), "composite_fwd")
        ms = timed(run, data)
        results.append({"key": key, "kernel": name, "rays": B, "S": s, "bytes_per_ray": bpr, "us": ms * 1e3,
                        "GBps": B * bpr / (ms / 1e3) / 1e9, "input_sets": len(data)})
        del data

    for key, name, with_u in (("sample_fine", f"sample_fine_fast_kernel<{Nc},{Nf},sorted> (in-kernel uniforms)", False),
                              ("sample_fine_u", f"sample_fine_fast_kernel<{Nc},{Nf}> (explicit uniforms)", True)):
        # weights + bin edges + t_coarse read, t_sorted written (+ u read)
        bpr = 4 * Nc * 2 + 4 * (Nc + 1) + 4 * S + (4 * Nf if with_u else 0)
        data = []
        for i in range(nsets(B * bpr)):
            t_c, edges = ru.sample_coarse(Nc, True, True, near, far, None, seed=i, ray0=0)
            data.append((ru.compute_weights(sigma_like(B * Nc), t_c), edges, t_c, rnd(B, Nf) if with_u else None))
        tso = torch.empty((B, S), device="cuda")

        def run(x):
            check(lib.nerfb200_sample_fine(B, Nc, Nf, ptr(x[0]), ptr(x[1]), ptr(x[2]), ptr(x[3], allow_none=True), 1, None, 0,
                                           ptr(tso), None, None, None, st), "sample_fine")
        ms = timed(run, data)
        results.append({"key": key, "kernel": name, "rays": B, "Nc": Nc, "Nf": Nf, "bytes_per_ray": bpr, "us": ms * 1e3,
                        "GBps": B * bpr / (ms / 1e3) / 1e9, "input_sets": len(data)})
        del data

    # stratified sampler: near/far read, t and the bin edges written, uniforms generated in-kernel
    bpr = 8 + 4 * Nc + 4 * (Nc + 1)
    outs = [(torch.empty((B, Nc), device="cuda"), torch.empty((B, Nc + 1), device="cuda")) for _ in range(nsets(B * bpr))]

    def run(x):
        check(lib.nerfb200_sample_coarse(B, Nc, 1, 1, ptr(near), ptr(far), None, 1, None, 0, ptr(x[0]), ptr(x[1]), st), "sample_coarse")
    ms = timed(run, outs)
    results.append({"key": "sample_coarse", "kernel": "sample_coarse_kernel (in-kernel uniforms, 1/t spacing)", "rays": B, "Nc": Nc,
                    "bytes_per_ray": bpr, "us": ms * 1e3, "GBps": B * bpr / (ms / 1e3) / 1e9, "input_sets": len(outs)})
    return results


def ncu_traffic(csv_name, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_substr`, averaged over the launches in a
    committed `ncu -i <rep> --page raw --csv` export under profiles/ (an `ncu --set full` capture). None if absent."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    if not os.path.exists(path):
        return None, None
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None, None
    hdr, units = rows[0], rows[1]
    try:
        kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    except ValueError:
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[rd].replace(",", "")) * scale.get(units[rd], 1.0) + float(r[wr].replace(",", "")) * scale.get(units[wr], 1.0)
            for r in rows[2:] if len(r) == len(hdr) and kernel_substr in r[kn]]
    if not vals:
        return None, None
    return sum(vals) / len(vals), f"profiles/{csv_name}: {len(vals)} launch(es) of {kernel_substr}, ncu --set full"


def run_b200(args):
    import torch
    import torch.distributed as dist
    import nerf_tf2_b200 as nb
    from nerf_tf2_b200 import ray_utils as ru, _lib, render as nbrender

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner (printed at every level >= VERSION) and
        # anything else it logs go to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL honours NCCL_DEBUG_FILE only above the VERSION level
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        return nb.dist.max_over_ranks(x, dev) if world > 1 else x

    lib = _lib.load()

    def timed(fn, steps, warmup):
        """W warm-up calls, then K calls between barrier + synchronize, CUDA events on the launching stream, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.nerfb200_launch_count()
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        timed.launches = lib.nerfb200_launch_count() - n0      # this library's kernels launched inside the timed region
        return max_ranks(e0.elapsed_time(e1)), t0, time.time()

    pk = peaks()
    tensor_peak = pk["tensor"] * (0.5 if args.precision == "tf32" else 1.0)
    tensor_burst = pk["tensor_burst"] * (0.5 if args.precision == "tf32" else 1.0)
    params = nb.make_params({"system": {"white_bg": True}}, N_coarse=N_COARSE, N_fine=N_FINE, perturb=True, lin_inv_depth=True)
    nerf = nb.setup_model(params, precision=args.precision, seed=0, rng_seed=1234, render_chunk=args.chunk)
    scene = nb.scene.SyntheticScene(H, W, num_cameras=max(8, world))
    n_rays = H * W
    warmup = max(args.warmup, 3)

    # ---- headline (`value`): ONE 800x800 view per step, its rays sharded over the ranks (contiguous ranges), replicated
    # weights, no data-path collective, and the final gather of the [H*W,5] image rows to rank 0 INSIDE the timed region
    a0, b0 = nb.dist.shard_range(n_rays, rank, world)
    shard = nb.create_dataset_for_render(H, W, scene.poses[0], scene.bounds, scene.K, batch_size=4096, on_device=True,
                                         ray0=a0, n_rays=b0 - a0).inputs
    keep = {}

    def step_view():
        keep["rows"] = nbrender.render_view_sharded(nerf, H, W, scene.poses[0], scene.bounds, scene.K, rays=shard, dst=0)

    clocks = ClockSampler(local) if rank == 0 else None
    ms, tw0, tw1 = timed(step_view, args.steps, warmup)
    launches_timed = timed.launches
    clk = clocks.stop(tw0, tw1) if clocks else None
    value = n_rays * args.steps / (ms / 1e3)
    gather_bytes = n_rays * 5 * 4 if world > 1 else 0

    # ---- the same K steps again with CUDA-event brackets around every C-ABI launch (on the launching stream)
    # -> per-kernel durations for the roofline objects (this rank's shard)
    kt = KernelTimer(torch)
    nerf.fused_forward = False        # this pass launches kernel by kernel from Python so that each can be bracketed
    orig = (nerf._mlp, ru.post_process_model_output, ru.sample_fine, ru.sample_coarse)
    nerf._mlp = kt.wrap("mlp", nerf._mlp)
    ru.post_process_model_output = kt.wrap("composite", ru.post_process_model_output)
    ru.sample_fine = kt.wrap("sample_fine", ru.sample_fine)
    ru.sample_coarse = kt.wrap("sample_coarse", ru.sample_coarse)
    ms_b, _, _ = timed(step_view, args.steps, 0)
    nerf._mlp, ru.post_process_model_output, ru.sample_fine, ru.sample_coarse = orig
    nerf.fused_forward = True
    my_rays = b0 - a0

    tot = kt.totals()
    mlp_ms, mlp_n = tot["mlp"]
    mlp_flop = args.steps * my_rays * ROWS_PER_RAY * FLOP_PER_ROW
    mlp_tflops = mlp_flop / (mlp_ms / 1e3) / 1e12
    comp_ms, comp_n = tot["composite"]
    # coarse: 24*S+20 B/ray with the [B,S] weights store; fine in render mode skips it: 20*S+20
    comp_bytes = args.steps * my_rays * ((24 * N_COARSE + 20) + (20 * (N_COARSE + N_FINE) + 20))
    sf_ms, sf_n = tot["sample_fine"]
    # weights 4Nc + bin edges 4(Nc+1) + t_coarse 4Nc read, t_sorted 4(Nc+Nf) written; u generated in-kernel
    sf_bytes = args.steps * my_rays * (4 * N_COARSE * 2 + 4 * (N_COARSE + 1) + 4 * (N_COARSE + N_FINE))
    sc_ms, sc_n = tot["sample_coarse"]
    # near/far read, t [B,Nc] and bin edges [B,Nc+1] written; u generated in-kernel
    sc_bytes = args.steps * my_rays * (8 + 4 * N_COARSE + 4 * (N_COARSE + 1))
    mlp_kernel = "mlp_tf32_forward_kernel" if args.precision == "tf32" else "mlp_tc_forward_pair_kernel"
    # (the main launch, not the small split launch whose name differs in the last template argument only)
    traffic, traffic_src = ncu_traffic(NCU_MLP_CSV.get(args.precision, ""),
                                       mlp_kernel if args.precision == "tf32" else "mlp_tc_forward_pair_kernel<0, 0, 0, 0>")
    roofline = {"bound": "tensor",
                "kernel": f"{mlp_kernel} (fused encoding + 8x256 MLP on tcgen05 cta_group::2; coarse, fine and split last-sample launches)",
                "achieved": mlp_tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": mlp_tflops / tensor_peak,
                "peak_kind": (f"bf16 sustained, {pk['source']}" if args.precision != "tf32"
                              else f"half of bf16 sustained ({pk['source']}): kind::tf32 runs at half the 16-bit rate"),
                "frac_of_burst": mlp_tflops / tensor_burst,
                "traffic": traffic, "traffic_source": traffic_src, "traffic_scope": "one fine launch of 65536 rays x 192 samples",
                "launches": mlp_n, "avg_launch_ms": mlp_ms / mlp_n, "share_of_step": mlp_ms / ms_b}
    in_step = {"composite": comp_bytes / (comp_ms / 1e3) / 1e9, "sample_fine": sf_bytes / (sf_ms / 1e3) / 1e9,
               "sample_coarse": sc_bytes / (sc_ms / 1e3) / 1e9}
    roofline_hbm = None
    if rank == 0 and not args.no_hbm:
        alone = {r["key"]: r for r in hbm_kernels_alone(rays=args.chunk)}
        # the integrator as the render uses it: one coarse launch (weights out) + one fine launch (no weights) per chunk
        ca, cf = alone["composite_coarse"], alone["composite_fine"]
        comp_alone = (ca["bytes_per_ray"] + cf["bytes_per_ray"]) * ca["rays"] / ((ca["us"] + cf["us"]) * 1e-6) / 1e9
        how = ("achieved: back-to-back launches of the kernel alone at the render chunk size, CUDA events around the batch, inputs "
               "rotated through sets > 2x L2; in_step_GBps: CUDA-event brackets around each single launch inside the render step "
               "(includes the launch gaps, which are comparable to a 15-50 us kernel)")
        tr_cc, src_h = ncu_traffic(NCU_HBM_CSV, "composite_fwd_kernel<4")
        tr_cf, _ = ncu_traffic(NCU_HBM_CSV, "composite_fwd_kernel<6")
        tr_sf, _ = ncu_traffic(NCU_HBM_CSV, "sample_fine_fast_kernel")
        roofline_hbm = [
            {"kernel": "composite_fwd_kernel<4,full,2 rays/warp> + <6,full> (coarse with weights, fine without)", "bound": "hbm",
             "achieved": comp_alone, "peak": pk["hbm"], "unit": "GB/s", "frac": comp_alone / pk["hbm"],
             "traffic": [tr_cc, tr_cf], "traffic_source": src_h,
             "us_per_launch": [ca["us"], cf["us"]], "frac_coarse": ca["GBps"] / pk["hbm"], "frac_fine": cf["GBps"] / pk["hbm"],
             "in_step_GBps": in_step["composite"], "launches": comp_n, "share_of_step": comp_ms / ms_b, "timing": how},
            {"kernel": alone["sample_fine"]["kernel"], "bound": "hbm", "achieved": alone["sample_fine"]["GBps"], "peak": pk["hbm"],
             "unit": "GB/s", "frac": alone["sample_fine"]["GBps"] / pk["hbm"], "traffic": tr_sf, "traffic_source": src_h,
             "us_per_launch": alone["sample_fine"]["us"], "in_step_GBps": in_step["sample_fine"], "launches": sf_n,
             "share_of_step": sf_ms / ms_b},
            {"kernel": alone["sample_coarse"]["kernel"], "bound": "hbm", "achieved": alone["sample_coarse"]["GBps"], "peak": pk["hbm"],
             "unit": "GB/s", "frac": alone["sample_coarse"]["GBps"] / pk["hbm"],
             "traffic": ncu_traffic(NCU_HBM_CSV, "sample_coarse_kernel")[0], "traffic_source": src_h,
             "us_per_launch": alone["sample_coarse"]["us"], "in_step_GBps": in_step["sample_coarse"], "launches": sc_n,
             "share_of_step": sc_ms / ms_b},
        ]
    barrier()

    # ---- weak scaling (secondary, N > 1): one WHOLE view per rank, no collective (round 1's headline)
    weak = None
    if world > 1:
        ds_dev = nb.create_dataset_for_render(H, W, scene.poses[rank % len(scene)], scene.bounds, scene.K, batch_size=4096, on_device=True)
        ro, rd, near, far = ds_dev.inputs
        ms_w, _, _ = timed(lambda: nerf.render_rays(ro, rd, near, far, need_weights=False), args.steps, 1)
        weak = {"value": world * n_rays * args.steps / (ms_w / 1e3), "unit": "rays/s", "ms_per_step": ms_w / args.steps,
                "what": "one whole 800x800 view per rank, no collective in the timed region"}
        del ds_dev, ro, rd, near, far

    # ---- end to end: HOST rays (pinned staging inside the call), H2D + D2H inside the timed region.
    # N = 1: the reference-facing call NeRF.predict(dataset). N > 1: every rank predicts its ray shard from host rays,
    # the device results are gathered to rank 0 (same collective as above) and copied to the host there.
    e2e = {"value": None, "unit": "rays/s", "h2d_bytes_per_step": n_rays * 32, "d2h_bytes_per_step": None}
    if not args.no_e2e:
        host = tuple(a.cpu().numpy() for a in shard)
        ds_host = nb.RayDataset.from_tensor_slices((host,)).batch(4096, drop_remainder=False)
        if world == 1:
            fn = lambda: nerf.predict(ds_host, return_weights=False)
            e2e["api"] = "NeRF.predict(RayDataset of host NumPy rays, return_weights=False) -> (dict_CM, dict_FM) of NumPy arrays"
            e2e["d2h_bytes_per_step"] = n_rays * 5 * 4 * 2
        else:
            fn = lambda: nbrender.predict_view_sharded(nerf, ds_host, n_rays, dst=0)
            e2e["api"] = ("render.predict_view_sharded(host ray shard per rank) -> NeRF.predict on the shard, gather of the "
                          "[H*W,5] rows to rank 0, D2H there")
            e2e["d2h_bytes_per_step"] = n_rays * 5 * 4
        for _ in range(2):
            fn()
        barrier()
        t0 = time.time()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        dt = max_ranks(time.time() - t0)
        e2e["value"] = n_rays * args.steps / dt
        # the Keras-default return (per-sample `weights` [N,S] in both dicts: 0.66 GB per view), N = 1 only
        if world == 1 and not args.no_extra:
            nerf.predict(ds_host, return_weights=True)
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(2):
                nerf.predict(ds_host, return_weights=True)
            torch.cuda.synchronize()
            dtw = (time.time() - t0) / 2
            e2e["keras_default_return_weights"] = {"value": n_rays / dtw, "unit": "rays/s", "ms_per_view": dtw * 1e3,
                                                   "d2h_bytes_per_step": n_rays * (5 * 2 + N_COARSE + N_COARSE + N_FINE) * 4}
        del ds_host, host
    barrier()

    # ---- secondary metric: 4096-ray training step (coarse+fine fwd/bwd + Adam [+ all-reduce])
    train = None
    if not args.no_train:
        try:
            train = bench_train(nb, nerf, torch, dist, world, rank, dev, barrier, args)
        except Exception as ex:  # report, do not hide
            train = {"error": str(ex)[:300]}

    # ---- named sub-workloads: BASELINE.json configs[3] and configs[4]
    workloads = {}
    if not args.no_extra:
        try:
            workloads["cfg4_eval_200_views"] = bench_cfg4(nb, nbrender, nerf, torch, dist, world, rank, barrier, max_ranks, args)
        except Exception as ex:  # report, do not hide
            workloads["cfg4_eval_200_views"] = {"error": str(ex)[:300]}
        try:
            workloads["cfg5_1080p_128_256"] = bench_cfg5(nb, nbrender, torch, world, rank, timed, args)
        except Exception as ex:
            workloads["cfg5_1080p_128_256"] = {"error": str(ex)[:300]}
    try:
        barrier()
    except Exception as ex:      # a secondary workload left the device or the communicator unusable: the headline measured
        workloads["barrier_after_secondary_workloads"] = {"error": str(ex)[:300]}      # above is still reported

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        arm = OracleArm((100, 100))          # BASELINE.json configs[0]: the reference's own CPU-runnable case, in full
        dt = arm.step()
        cpu = {"value": arm.n / dt, "unit": "rays/s", "cores": arm.threads, "kind": "port", "seconds": dt,
               "sample": "BASELINE.json configs[0] in full: one 100x100 view (10 000 rays), coarse+fine 64+128, forward only, "
                         "fp32 torch-CPU/NumPy oracle (TensorFlow not installable)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": WORKLOAD, "rays_per_step": n_rays, "rays_per_step_per_gpu": my_rays,
                           "render_chunk_rays": args.chunk,
                           "l2": "inputs larger than L2 (per-chunk rgb/sigma/t buffers > 126 MB); no explicit flush",
                           "parallelism": (f"ONE view per step, rays sharded x{world} in contiguous ranges, replicated weights; "
                                           f"final all-gather of the [H*W,5] image rows ({gather_bytes} B) inside the timed region"
                                           if world > 1 else "single GPU, no collective"),
                           "split_last_sample_launch": True},
                "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches_timed), "clocks": clk, "weak": weak, "workloads": workloads, "train": train}
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing the communicator down: every rank has passed the last collective (barrier), the line is
        # out, and ncclCommDestroy at interpreter exit is where multi-rank jobs hang when anything still holds NCCL work
        import gc
        gc.collect()
        try:
            barrier()
        except Exception:
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def bench_cfg4(nb, nbrender, nerf, torch, dist, world, rank, barrier, max_ranks, args):
    """BASELINE.json configs[3]: evaluate.py-shaped test-set render -- 200 views of 800x800 (main/eval.py:42-67 with the
    depth maps of main/render.py:96-117): per view ray generation, ray march, uint8 image, PSNR against the ground-truth
    image, depth map types 1 and 2; views dealt round-robin to the ranks, ONE all-reduce of the per-view squared errors
    at the end. Ground truth: one synthetic uint8 image reused for every view (the PSNR value is not the point here)."""
    n_views = args.cfg4_views
    poses = nb.scene.SyntheticScene(H, W, num_cameras=n_views)
    gt = np.random.default_rng(0).integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    gts = [gt] * n_views
    run = lambda p, g: nbrender.evaluate_views(nerf, H, W, p, poses.bounds, poses.K, g, scale_factor=poses.adj_scale_factor,
                                               depth_maps=True)
    run(poses.poses[:world], gts[:world])       # warm-up: one view per rank
    barrier()
    t0 = time.time()
    res = run(poses.poses, gts)
    torch.cuda.synchronize()
    dt = max_ranks(time.time() - t0)
    rays = n_views * H * W
    return {"what": f"{n_views} views of 800x800, 64+128 samples: rays -> march -> uint8 + PSNR + depth type_1/type_2, "
                    f"view-sharded x{world}, final PSNR all-reduce",
            "seconds": dt, "views_per_s": n_views / dt, "value": rays / dt, "unit": "rays/s",
            "mlp_tflops": rays * ROWS_PER_RAY * FLOP_PER_ROW / dt / 1e12, "mean_psnr": res["mean_psnr"]}


def bench_cfg5(nb, nbrender, torch, world, rank, timed, args):
    """BASELINE.json configs[4]: 1920x1080 render with 128+256 samples, rays sharded over the ranks, final image gather."""
    h5, w5, nc5, nf5 = 1080, 1920, 128, 256
    p5 = nb.make_params({"system": {"white_bg": True}}, N_coarse=nc5, N_fine=nf5, perturb=True, lin_inv_depth=True)
    n5 = nb.setup_model(p5, precision=args.precision, seed=0, rng_seed=99, render_chunk=args.chunk)
    sc5 = nb.scene.SyntheticScene(h5, w5, num_cameras=8)
    a, b = nb.dist.shard_range(h5 * w5, rank, world)
    sh = nb.create_dataset_for_render(h5, w5, sc5.poses[0], sc5.bounds, sc5.K, batch_size=4096, on_device=True,
                                      ray0=a, n_rays=b - a).inputs
    keep = {}

    def step():
        keep["rows"] = nbrender.render_view_sharded(n5, h5, w5, sc5.poses[0], sc5.bounds, sc5.K, rays=sh, dst=0)

    steps = 3
    ms, _, _ = timed(step, steps, 1)
    rays = h5 * w5
    rows = nc5 + nc5 + nf5
    return {"what": f"1920x1080 view, 128+256 samples, rays sharded x{world}, final gather of the [H*W,5] rows inside the timed region",
            "ms_per_view": ms / steps, "value": rays * steps / (ms / 1e3), "unit": "rays/s",
            "mlp_tflops": rays * rows * FLOP_PER_ROW * steps / (ms / 1e3) / 1e12}


def bench_train(nb, nerf, torch, dist, world, rank, dev, barrier, args):
    B = 4096
    Bl = B // world
    scene = nb.scene.SyntheticScene(H, W, num_cameras=8)
    g = torch.Generator(device="cpu").manual_seed(rank)
    ids = torch.randint(0, H * W, (Bl,), generator=g, dtype=torch.int32).to(dev)
    ro, rd = nb.ray_utils.get_rays_at(H, W, scene.K, scene.poses[0], ids)
    near = torch.full((Bl, 1), scene.near, device=dev); far = torch.full((Bl, 1), scene.far, device=dev)
    rgb = torch.rand((Bl, 3), device=dev)
    p = nb.make_params({"system": {"white_bg": True}})
    tp = args.train_precision
    batch = ((ro, rd, near, far), (rgb,))
    steps = 50

    def run(graph, peer=True):
        tn = nb.setup_model(p, precision=tp if tp != "fp32" else "bf16", train_precision=tp, seed=0, cuda_graph=graph,
                            precise_last=False)
        if world > 1:
            tn.set_distributed(peer_exchange=peer)
        used_peer = tn.peer_mode
        for _ in range(5):
            tn.train_step(batch)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tn.train_step(batch)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            ms = nb.dist.max_over_ranks(ms, dev)
        captured = any("graph" in st for st in tn._graphs.values())
        tn.release_cuda_graphs()      # captured NCCL work must be gone before the process group is torn down
        tn.close_distributed()        # unmaps the peers' gradient buffers (collective)
        del tn
        return ms, captured, used_peer

    ms_eager, _, used_peer = run(False)
    ms_graph, captured, _ = run(True) if not args.no_train_graph else (ms_eager, False, used_peer)
    ms = min(ms_eager, ms_graph) if captured else ms_eager      # `value`: the faster of the two launch modes
    sps = steps / (ms / 1e3)
    nccl = None
    if world > 1 and used_peer:   # the same step with the gradient handed to NCCL instead (what the peer kernel replaces)
        ms_n, cap_n, _ = run(not args.no_train_graph, peer=False)
        nccl = {"value": steps / (ms_n / 1e3), "ms_per_step": ms_n / steps, "cuda_graph": bool(cap_n)}
    flop = 3489024 * B * ROWS_PER_RAY
    pk = peaks()
    return {"metric": "train steps/s (4096-ray batch, coarse+fine fwd/bwd + Adam, data-parallel all-reduce)",
            "value": sps, "unit": "steps/s", "ms_per_step": ms / steps, "steps": steps, "precision": tp,
            "global_batch": B, "rays_per_gpu": Bl, "achieved_tflops": flop * sps / 1e12,
            "gradient_exchange": (None if world == 1 else "NCCL all-reduce" if not used_peer else
                                  {"kernel": "one launch per rank over NVLink peer memory, fused with Adam (csrc/peer.cu)",
                                   "mapping": used_peer,
                                   "data_path": ("NVLS: multimem.ld_reduce / multimem.st through the NVSwitch multicast mapping"
                                                 if used_peer == "nvls" else "unicast peer loads / stores (reduce-scatter + all-gather)")}),
            "with_nccl_allreduce_instead": nccl,
            "frac_of_sustained_peak": flop * sps / 1e12 / (pk["tensor"] * world),
            "launch_mode": ("one CUDA graph per step (sampling, forwards, loss, backwards, all-reduce, Adam, repack)"
                            if captured and ms_graph <= ms_eager else "eager launches"),
            "eager": {"value": steps / (ms_eager / 1e3), "ms_per_step": ms_eager / steps},
            "cuda_graph": {"value": steps / (ms_graph / 1e3), "ms_per_step": ms_graph / steps} if captured else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "tf32"],
                    help="MLP operand width of the render (tf32 = the reference's own width on Ampere-and-later GPUs)")
    ap.add_argument("--train-precision", default="bf16", choices=["fp32", "bf16", "fp16"])
    ap.add_argument("--chunk", type=int, default=65536, help="rays per internal render chunk")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--no-hbm", action="store_true", help="skip the stand-alone timings of the HBM-bound kernels")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg4 / cfg5 sub-workloads and the return_weights=True leg")
    ap.add_argument("--cfg4-views", type=int, default=200)
    ap.add_argument("--no-train-graph", action="store_true", help="time the training step with eager launches only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
