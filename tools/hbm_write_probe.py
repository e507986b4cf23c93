import torch
x = torch.empty(4_270_000_000, dtype=torch.uint8, device="cuda")
y = torch.empty_like(x)
for name, fn in (("fill (write only)", lambda: x.fill_(1)), ("copy (read+write)", lambda: y.copy_(x))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms -> {x.numel()/ms/1e6:.1f} GB/s of payload")
