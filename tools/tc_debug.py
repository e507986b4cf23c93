"""Developer tool: prints the in-kernel cycle counters of the fused MLP kernel (debug instantiation, selected with
nerfb200_set_option(ctx, NERFB200_OPT_DEBUG, 1))."""
import sys
sys.path.insert(0, ".")
import torch
import nerf_tf2_b200 as nb
from nerf_tf2_b200 import _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
nerf = nb.setup_model(nb.make_params(), precision=prec)
_lib.check(_lib.load().nerfb200_set_option(nerf._ctx, _lib.OPT_DEBUG, 1), "set_option")
B, S = 65536, 192
ro = torch.zeros((B, 3), device="cuda"); rd = torch.nn.functional.normalize(torch.randn((B, 3), device="cuda"), dim=1)
t = torch.sort(torch.rand((B, S), device="cuda") * 0.85 + 0.425, dim=1)[0].contiguous()
for i in range(2):
    nerf._mlp(1, ro, rd, t, _lib.BF16 if prec == "bf16" else _lib.FP16)
    torch.cuda.synchronize()
