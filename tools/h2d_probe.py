import time, torch
dev = torch.device("cuda:0")
h = torch.rand(1, 1080, 1920, 3).pin_memory()
d = torch.empty_like(h, device=dev)
print("pinned:", h.is_pinned())
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"H2D 24.9 MB: {dt*1e3:.3f} ms -> {h.numel()*4/dt/1e9:.2f} GB/s")
x = torch.zeros(1, device=dev)
t0 = time.perf_counter()
for _ in range(100):
    x.item()
print(f".item(): {(time.perf_counter()-t0)/100*1e6:.1f} us")
