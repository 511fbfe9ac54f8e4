"""Probe: cost of uploading an 8-bit image and converting it on the device (bench.py's e2e input path)."""
import torch

dev = torch.device("cuda:0")
u8 = torch.randint(0, 256, (1, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
f32 = (u8.float() / 255).pin_memory()
d_u8 = torch.empty_like(u8, device=dev)
d_f = torch.empty((1, 1080, 1920, 3), device=dev)
print("pinned:", u8.is_pinned(), f32.is_pinned(), d_u8.device, d_u8.dtype)


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


print("h2d u8   %.1f us" % t(lambda: d_u8.copy_(u8, non_blocking=True)))
print("h2d f32  %.1f us" % t(lambda: d_f.copy_(f32, non_blocking=True)))
print("mul out  %.1f us" % t(lambda: torch.mul(d_u8, 1.0 / 255.0, out=d_f)))
print("float+mul_ %.1f us" % t(lambda: d_u8.float().mul_(1.0 / 255.0)))
print("copy_ cast %.1f us" % t(lambda: d_f.copy_(d_u8)))
