"""Developer diagnostics: gradients of the tensor-core training path vs the on-device fp32 path."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import nerf_tf2_b200 as nb
from oracle import scene as osc

def grads(precision, B, seed=0):
    v = osc.synthetic_view(32, 32, view=1)
    rng = np.random.default_rng(seed)
    sel = rng.choice(1024, size=B, replace=B > 1024)
    gt = rng.random((B, 3), dtype=np.float32)
    uf = torch.from_numpy(rng.random((B, 128), dtype=np.float32)).cuda()
    nerf = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=False), precision=precision, seed=3)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    loss, pc, pf = nerf._loss_and_grads(dev(v["rays_o"][sel]), dev(v["rays_d"][sel]), dev(v["near"][sel]), dev(v["far"][sel]), dev(gt), u_fine=uf)
    torch.cuda.synchronize()
    return nerf, float(loss.item()), nerf.flat_grads.clone()

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n32, l32, g32 = grads("fp32", B)
for prec in sys.argv[2:] or ["bf16"]:
    n16, l16, g16 = grads(prec, B)
    print(f"B={B} loss fp32 {l32:.6f} {prec} {l16:.6f} rel {abs(l16-l32)/l32:.2e}")
    cos = torch.nn.functional.cosine_similarity(g32.double(), g16.double(), dim=0).item()
    print(f"global cosine {cos:.6f} rel-L2 {(g16-g32).norm().item()/g32.norm().item():.3e} |g32| {g32.norm().item():.4e} |g16| {g16.norm().item():.4e} nan={int(torch.isnan(g16).sum())}")
    for var in n32.trainable_variables:
        a = g32[var._ofs:var._ofs + var._n].double(); b = g16[var._ofs:var._ofs + var._n].double()
        c = torch.nn.functional.cosine_similarity(a, b, dim=0).item() if a.norm() > 0 and b.norm() > 0 else float("nan")
        print(f"  {var.name:28s} |ref| {a.norm().item():.3e} |tc| {b.norm().item():.3e} cos {c:.5f} relL2 {(a-b).norm().item()/(a.norm().item()+1e-30):.3e}")
