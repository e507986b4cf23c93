"""Developer tool: N tensor-core train steps on a B-ray batch (for ncu launch lists).
usage: python tools/train_bench.py [steps] [precision] [B] [graph]     (B = 512 is one rank's share of a data-parallel step on 8 GPUs)"""
import sys
import torch
sys.path.insert(0, ".")
import nerf_tf2_b200 as nb
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
sc = nb.scene.SyntheticScene(800, 800)
ids = torch.randint(0, 640000, (B,), dtype=torch.int32).cuda()
ro, rd = nb.ray_utils.get_rays_at(800, 800, sc.K, sc.poses[0], ids)
near = torch.full((B, 1), sc.near, device="cuda"); far = torch.full((B, 1), sc.far, device="cuda")
rgb = torch.rand((B, 3), device="cuda")
nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}), precision=prec, train_precision=prec, cuda_graph=graph,
                      precise_last=False)
batch = ((ro, rd, near, far), (rgb,))
for _ in range(steps):
    nerf.train_step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    nerf.train_step(batch)
e1.record(); torch.cuda.synchronize()
print(f"{steps} steps of {B} rays ({'graph' if graph else 'eager'}): {e0.elapsed_time(e1)/steps:.3f} ms/step")
