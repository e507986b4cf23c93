"""Developer tool: two fused-MLP launches of the fine-pass shape (65536 rays x 192 samples), for ncu captures."""
import sys
sys.path.insert(0, ".")
import torch
import nerf_tf2_b200 as nb
from nerf_tf2_b200 import _lib
nerf = nb.setup_model(nb.make_params(), precision="bf16")
B, S = 65536, 192
ro = torch.zeros((B, 3), device="cuda"); rd = torch.nn.functional.normalize(torch.randn((B, 3), device="cuda"), dim=1)
t = torch.sort(torch.rand((B, S), device="cuda") * 0.85 + 0.425, dim=1)[0].contiguous()
for i in range(2):
    nerf._mlp(1, ro, rd, t, _lib.BF16)
    torch.cuda.synchronize()
print("done")
