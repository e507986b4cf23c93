"""Developer tool (torchrun, N GPUs): the data-parallel training step under different schedules --
eager / CUDA graph, the gradient exchange by the library's peer-memory kernel (fused with Adam or not) or by NCCL
(all-reduce of the coarse half overlapped with the fine backward or not). (Leaving 4-16 SMs of the
fine backward free for NCCL was measured too, profiles/r2k_dp_n8.json: no gain -- the option was removed again.)
usage: torchrun --nproc-per-node N tools/dp_bench.py [steps]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import nerf_tf2_b200 as nb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
only_exchange = len(sys.argv) > 2 and sys.argv[2] == "exchange"      # skip the schedules, time the exchange alone
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
B = 4096 // world
sc = nb.scene.SyntheticScene(800, 800)
g = torch.Generator().manual_seed(rank)
ids = torch.randint(0, 640000, (B,), generator=g, dtype=torch.int32).to(dev)
ro, rd = nb.ray_utils.get_rays_at(800, 800, sc.K, sc.poses[0], ids)
near = torch.full((B, 1), sc.near, device=dev); far = torch.full((B, 1), sc.far, device=dev)
rgb = torch.rand((B, 3), device=dev)
batch = ((ro, rd, near, far), (rgb,))
res = {}
for name, kw in (("nccl_eager_overlap", dict(graph=False, overlap=True, peer=False)),
                 ("nccl_graph", dict(graph=True, overlap=False, peer=False)),
                 ("nccl_graph_overlap", dict(graph=True, overlap=True, peer=False)),
                 ("peer_eager", dict(graph=False, peer=True)),
                 ("peer_graph_separate_adam", dict(graph=True, peer=True, fuse=False)),
                 ("peer_graph", dict(graph=True, peer=True)),
                 ("peer_graph_unicast", dict(graph=True, peer=True, multicast=False)),
                 ("peer_graph_ipc", dict(graph=True, peer=True, symm=False)),
                 ("no_exchange_graph", dict(graph=True, dist=False))):
    if only_exchange:
        break
    tn = nb.setup_model(nb.make_params({"system": {"white_bg": True}}), precision="bf16", train_precision="bf16", cuda_graph=kw["graph"],
                        precise_last=False)
    tn.peer_symmetric_memory, tn.peer_multicast = kw.get("symm", True), kw.get("multicast", None)
    if kw.get("dist", True):
        tn.set_distributed(peer_exchange=kw.get("peer", True))
    tn.overlap_allreduce = kw.get("overlap", True)
    tn.fuse_exchange_adam = kw.get("fuse", True)
    tn.train_step(batch)
    # after ONE step the schedules may differ by fp32 summation order only (NCCL reduces a buffer split in two in a
    # different rank order than the same buffer in one piece); later steps amplify that chaotically
    first = tn.flat_params.double().clone()
    if "ref_first" not in res:
        res["ref_first"] = first
    step1_diff = float((first - res["ref_first"]).abs().max())
    for _ in range(7):
        tn.train_step(batch)
    captured = any("graph" in st for st in tn._graphs.values())
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tn.train_step(batch)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = nb.dist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    res[name] = {"ms_per_step": ms, "steps_per_s": 1e3 / ms, "captured": captured, "loss": float(tn.last_loss.item()),
                 "param_sum": float(tn.flat_params.double().sum().item()), "max_param_diff_after_step_1_vs_first_schedule": step1_diff}
    res[name]["peer_exchange"] = tn.peer_mode
    tn.release_cuda_graphs()
    tn.close_distributed()
    del tn
res.pop("ref_first", None)

# ---- the exchange alone, back to back on the 4.77 MB gradient buffer: the library's peer kernel (NVLS multicast, unicast over
# symmetric memory, unicast over the library's own IPC mapping) vs NCCL
import ctypes as C
from nerf_tf2_b200 import _lib
lib = _lib.load()
n = _lib.PARAMS_TOTAL + 4
nccl_buf = torch.zeros(n, device=dev)
opt = [torch.zeros(n - 4, device=dev) for _ in range(3)]


def alone(fn, iters=300):
    for _ in range(20):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    return nb.dist.max_over_ranks(e0.elapsed_time(e1), dev) / iters * 1e3


exch = {"buffer_bytes": 4 * n}
for label, symm, mcast in (("multicast", True, True), ("unicast", True, False), ("ipc", False, False)):
    tn = nb.setup_model(nb.make_params({"system": {"white_bg": True}}), precision="bf16", train_precision="bf16", precise_last=False)
    tn.peer_symmetric_memory, tn.peer_multicast = symm, mcast
    tn.set_distributed()
    tn._grad_buf.zero_()
    h = tn._peer
    exch[label] = {"mode": tn.peer_mode,
                   "peer_kernel_us": alone(lambda: _lib.check(lib.nerfb200_peer_allreduce(h, _lib.stream_ptr()), "peer_allreduce")),
                   "peer_kernel_with_adam_us": alone(lambda: _lib.check(lib.nerfb200_peer_allreduce_adam(
                       h, n - 4, _lib.ptr(opt[0]), _lib.ptr(opt[1]), _lib.ptr(opt[2]), 0, None, _lib.stream_ptr()), "peer_allreduce_adam"))}
    # where the time goes inside ONE exchange (the kernel's own %globaltimer stamps, per rank): the last of 20 back to back
    for fused in (False, True):
        for _ in range(20):
            if fused:
                _lib.check(lib.nerfb200_peer_allreduce_adam(h, n - 4, _lib.ptr(opt[0]), _lib.ptr(opt[1]), _lib.ptr(opt[2]), 0, None,
                                                            _lib.stream_ptr()), "peer_allreduce_adam")
            else:
                _lib.check(lib.nerfb200_peer_allreduce(h, _lib.stream_ptr()), "peer_allreduce")
        torch.cuda.synchronize()
        st = (C.c_ulonglong * 7)()
        _lib.check(lib.nerfb200_peer_profile(h, st), "peer_profile")
        mine_t = torch.tensor([float(st[i] - st[0]) / 1e3 for i in (0, 1, 6, 2, 3, 4, 5)], device=dev)
        allt = [torch.empty_like(mine_t) for _ in range(world)]
        dist.all_gather(allt, mine_t)
        exch[label]["stamps_us_with_adam" if fused else "stamps_us"] = {
            "what": "per rank, us since kernel start: barrier A passed, loads back + stores issued, system fence done (first CTA); all CTAs done, barrier B passed, end (last CTA)",
            "ranks": [[round(float(x), 2) for x in t[1:].tolist()] for t in allt]}
    tn.close_distributed()
    del tn
exch["adam_kernel_alone_us"] = alone(lambda: _lib.check(lib.nerfb200_adam_step(
    n - 4, _lib.ptr(opt[0]), _lib.ptr(nccl_buf), _lib.ptr(opt[1]), _lib.ptr(opt[2]), 0, None, _lib.stream_ptr()), "adam_step"))
exch["nccl_all_reduce_us"] = alone(lambda: dist.all_reduce(nccl_buf))
if rank == 0:
    print(json.dumps({"world": world, "rays_per_gpu": B, "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"), "results": res,
                      "exchange_alone": exch}), flush=True)
dist.barrier()
sys.stdout.flush()
os._exit(0)
