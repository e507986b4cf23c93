"""HBM-bound kernels timed alone (the integrator, the hierarchical and the stratified sampler) at the render chunk
size: CLI around bench.hbm_kernels_alone(). Prints one JSON object per kernel: algorithmic bytes per ray
(SURVEY.md section 8d / DESIGN.md section 4), microseconds per launch, GB/s, fraction of the measured HBM peak.

    python tools/hbm_kernels_bench.py [--rays 65536] [--iters 40] [--out gpurun_out/hbm_kernels.json]
    ncu --set full -k 'regex:composite_fwd|sample_' -c 10 python tools/hbm_kernels_bench.py --iters 1 --warmup 0 --sets 1
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


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
    pk = bench.peaks()
    results = bench.hbm_kernels_alone(a.rays, a.iters, a.warmup, a.sets, a.nc, a.nf)
    for r in results:
        r["frac_of_hbm_peak"] = r["GBps"] / pk["hbm"]
        r["hbm_peak_GBps"] = pk["hbm"]
        r["peak_source"] = pk["source"]
        print(json.dumps(r))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
