"""Developer tool: two fused-MLP forwards of the fine-pass shape (65536 rays x 192 samples), for ncu captures.
usage: python tools/mlp_fine_launch.py [bf16|fp16|tf32]   (each forward = the main launch + the split last-sample launch)"""
import sys
sys.path.insert(0, ".")
import torch
import nerf_tf2_b200 as nb
from nerf_tf2_b200 import _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
nerf = nb.setup_model(nb.make_params(), precision=prec)
B, S = 65536, 192
ro = torch.zeros((B, 3), device="cuda"); rd = torch.nn.functional.normalize(torch.randn((B, 3), device="cuda"), dim=1)
t = torch.sort(torch.rand((B, S), device="cuda") * 0.85 + 0.425, dim=1)[0].contiguous()
for i in range(2):
    nerf._mlp(1, ro, rd, t, _lib.PRECISIONS[prec])
    torch.cuda.synchronize()
print("done")
