"""Device timings (CUDA events) of the section-8(f) kernels at the headline size: photometric loss at 1080p,
GroupedAdam over the six parameter tensors of 1 M Gaussians vs torch.optim.Adam, split-SH colour stage vs the
concatenated form.  Development aid; bench.py is the contract."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gscodec_studio_b200 import _cabi  # noqa: E402
from gscodec_studio_b200.cuda._wrapper import view_colors_from_sh, view_colors_from_sh_split  # noqa: E402
from gscodec_studio_b200.losses import photometric_loss  # noqa: E402
from gscodec_studio_b200.optimizers import GroupedAdam  # noqa: E402

DEV = torch.device("cuda:0")


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[len(ts) // 2] * 1e3  # median, us


def device_ms(fn, n=10):
    """Device time of every C-ABI call inside fn (CUDA events around the calls), averaged, in us."""
    fn()
    _cabi.enable_stage_timing(True)
    for _ in range(n):
        fn()
    ms = _cabi.stage_timing_ms()
    _cabi.enable_stage_timing(False)
    return ", ".join(f"{k.replace('gsb200_', '')} {v * 1e3:.1f} us" for k, v in sorted(ms.items()))


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    # ---- loss
    x = torch.rand(1, 1080, 1920, 3, device=DEV, generator=g).requires_grad_(True)
    y = torch.rand(1, 1080, 1920, 3, device=DEV, generator=g)

    def ours():
        x.grad = None
        photometric_loss(x, y, 0.2).backward()

    def torch_l1():
        x.grad = None
        (x - y).abs().mean().backward()

    print("  device:", device_ms(ours))
    print(f"photometric_loss (L1+SSIM) fwd+bwd 1080p: {timeit(ours):8.1f} us   | torch L1 only fwd+bwd: {timeit(torch_l1):8.1f} us")

    # ---- Adam
    N = 1_006_065
    shapes = {"means": (N, 3), "scales": (N, 3), "quats": (N, 4), "opacities": (N,), "sh0": (N, 1, 3), "shN": (N, 15, 3)}
    lrs = {"means": 1.6e-4, "scales": 5e-3, "quats": 1e-3, "opacities": 5e-2, "sh0": 2.5e-3, "shN": 2.5e-3 / 20}

    def build(cls, **kw):
        ps = {k: torch.nn.Parameter(torch.randn(*s, device=DEV, generator=g)) for k, s in shapes.items()}
        for p in ps.values():
            p.grad = torch.randn_like(p)
        return ps, cls([{"params": [p], "lr": lrs[k], "name": k} for k, p in ps.items()], eps=1e-15, betas=(0.9, 0.999), **kw)

    _, grouped = build(GroupedAdam)
    _, foreach = build(torch.optim.Adam)
    _, fused = build(torch.optim.Adam, fused=True)
    n_bytes = 28 * sum(int(torch.tensor(s).prod()) for s in shapes.values())
    print("  device:", device_ms(grouped.step))
    t = timeit(grouped.step)
    print(f"GroupedAdam 6 tensors / 1 launch: {t:8.1f} us ({n_bytes / t / 1e3:7.1f} GB/s of 28 B/float) | "
          f"torch Adam foreach (one optimizer): {timeit(foreach.step):8.1f} us | torch Adam fused=True: {timeit(fused.step):8.1f} us")

    # ---- split SH
    means = torch.randn(N, 3, device=DEV, generator=g)
    campos = torch.tensor([[0.0, 0.0, 5.0]], device=DEV)
    sh0 = torch.randn(N, 1, 3, device=DEV, generator=g).requires_grad_(True)
    shN = torch.randn(N, 15, 3, device=DEV, generator=g).requires_grad_(True)
    radii = ((torch.rand(1, N, device=DEV, generator=g) < 0.29).int() * 4).contiguous()
    w = torch.randn(1, N, 3, device=DEV, generator=g)

    def split():
        sh0.grad = shN.grad = None
        (view_colors_from_sh_split(3, means, campos, sh0, shN, radii) * w).sum().backward()

    def cat():
        sh0.grad = shN.grad = None
        (view_colors_from_sh(3, means, campos, torch.cat([sh0, shN], 1), radii) * w).sum().backward()

    print("  device:", device_ms(split), "|", device_ms(cat))
    print(f"colour stage fwd+bwd, 29% visible: split (sh0, shN) {timeit(split):8.1f} us | torch.cat + fused kernel {timeit(cat):8.1f} us")


if __name__ == "__main__":
    main()
