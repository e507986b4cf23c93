"""Developer tool: time the fused forward with and without the training stash (fine-model shape of a 4096-ray step).
With a mode argument 3|4|5 (nerfb200_set_option NERFB200_OPT_DEBUG) the debug instantiation drops the stash stores / the ReLU bitmask stores / both, which
attributes the training forward's slowdown over the inference forward (results are then meaningless)."""
import os
import sys
import torch
sys.path.insert(0, ".")
import nerf_tf2_b200 as nb
from nerf_tf2_b200 import _lib
nerf = nb.setup_model(nb.make_params(), precision="bf16")
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if mode:
    _lib.check(_lib.load().nerfb200_set_option(nerf._ctx, _lib.OPT_DEBUG, mode), "set_option")
B, S = 4096, 192
ro = torch.zeros((B, 3), device="cuda"); rd = torch.nn.functional.normalize(torch.randn((B, 3), device="cuda"), dim=1)
t = torch.sort(torch.rand((B, S), device="cuda") * 0.85 + 0.425, dim=1)[0].contiguous()
stash = torch.empty(_lib.load().nerfb200_mlp_stash_bytes(B * S, _lib.BF16), device="cuda", dtype=torch.uint8)
def run(st, n=10):
    for _ in range(3): nerf._mlp(1, ro, rd, t, _lib.BF16, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): nerf._mlp(1, ro, rd, t, _lib.BF16, st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"debug mode {mode} rows {B*S}: inference {run(None):.3f} ms, "
      f"training (stash {stash.numel()/1e9:.2f} GB) {run(stash):.3f} ms")
