"""GPU box: host<->device copy bandwidth from pinned memory (what bounds the e2e path's H2D/D2H legs)."""
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(src, non_blocking=True); e1.record(); e1.synchronize()
        best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    print(f"{name}: {best:.1f} GB/s (1 GiB, pinned)")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); e1.record(); e1.synchronize()
print(f"H2D+D2H concurrent: {2 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s total")
