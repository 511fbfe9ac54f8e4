"""GPU parity tests of the rows SURVEY.md section 8(f) marks "next" and of the additive fusion hooks:
split SH coefficients, the fused optimizer step, the un-fused public ops.  Same conventions as
tests/test_gpu_parity.py (tolerances relative to the largest reference magnitude)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import gscodec_studio_b200 as gs  # noqa: E402
from gscodec_studio_b200._helper import load_test_data  # noqa: E402
from oracle import oracle as O  # noqa: E402

DEV = torch.device("cuda:0")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


# ------------------------------------------------------------------------------- split SH coefficients
@pytest.mark.parametrize("deg,K,N,C", [(0, 1, 257, 1), (0, 16, 1000, 2), (1, 4, 130, 3), (2, 9, 3001, 1), (3, 16, 3001, 3),
                                       (1, 16, 515, 1), (4, 25, 1029, 2), (3, 16, 128, 1), (3, 16, 5, 1)])
def test_split_sh_matches_concatenated(deg, K, N, C):
    """view_colors_from_sh_split(sh0, shN) == view_colors_from_sh(cat(sh0, shN)) (same per-Gaussian
    arithmetic), and the gradients of sh0 / shN / means / campos match the concatenated
    form (which test_gpu_parity pins against the oracle); rows of invisible Gaussians and unused bands get
    exact zeros."""
    from gscodec_studio_b200.cuda._wrapper import view_colors_from_sh, view_colors_from_sh_split

    g = torch.Generator().manual_seed(100 * deg + K)
    means = torch.randn(N, 3, generator=g)
    campos = torch.randn(C, 3, generator=g) * 3
    coeffs = torch.randn(N, K, 3, generator=g) * 0.7
    radii = ((torch.rand(C, N, generator=g) > 0.6).int() * 5).to(DEV)
    w = torch.randn(C, N, 3, generator=g).to(DEV)
    m1, c1 = means.to(DEV).requires_grad_(True), campos.to(DEV).requires_grad_(True)
    sh0 = coeffs[:, :1].contiguous().to(DEV).requires_grad_(True)
    shN = coeffs[:, 1:].contiguous().to(DEV).requires_grad_(True)
    m2, c2 = means.to(DEV).requires_grad_(True), campos.to(DEV).requires_grad_(True)
    full = coeffs.to(DEV).requires_grad_(True)

    out_split = view_colors_from_sh_split(deg, m1, c1, sh0, shN, radii)
    out_full = view_colors_from_sh(deg, m2, c2, full, radii)
    assert rel_err(out_split, out_full) < 1e-6
    (out_split * w).sum().backward()
    (out_full * w).sum().backward()
    assert rel_err(sh0.grad, full.grad[:, :1]) < 1e-6
    if K > 1:
        assert rel_err(shN.grad, full.grad[:, 1:]) < 1e-6
        invisible = (radii <= 0).all(dim=0)
        assert (shN.grad[invisible] == 0).all()
        nb = (deg + 1) ** 2
        assert (shN.grad[:, nb - 1:] == 0).all()
    assert rel_err(m1.grad, m2.grad) < 1e-6
    assert rel_err(c1.grad, c2.grad) < 1e-4  # block-level atomics: summation order differs


@pytest.mark.parametrize("packed", [False, True])
def test_rasterization_accepts_sh_pair(packed):
    """rasterization(colors=(sh0, shN)) renders and differentiates exactly like colors=cat(sh0, shN)."""
    means, quats, scales, opac, colors, viewmats, Ks, W, H = load_test_data(device="cpu", scene_grid=1)
    sel = slice(0, None, 5)
    g = torch.Generator().manual_seed(5)
    N = means[sel].shape[0]
    base = [means[sel], quats[sel], scales[sel], opac[sel]]
    sh0 = ((colors[sel] - 0.5) / 0.2820947917738781)[:, None, :]
    shN = torch.randn(N, 15, 3, generator=g) * 0.1
    viewmats, Ks = viewmats[:2].to(DEV), Ks[:2].to(DEV)

    def run(pair: bool):
        p = [t.clone().to(DEV).requires_grad_(True) for t in base]
        a, b = sh0.clone().to(DEV).requires_grad_(True), shN.clone().to(DEV).requires_grad_(True)
        cols = (a, b) if pair else torch.cat([a, b], 1)
        rc, ra, meta = gs.rasterization(p[0], p[1], p[2], p[3], cols, viewmats, Ks, W, H, sh_degree=3, packed=packed)
        (rc.sum() + 0.3 * ra.sum()).backward()
        return rc, ra, [t.grad for t in p] + [a.grad, b.grad]

    rc1, ra1, g1 = run(True)
    rc2, ra2, g2 = run(False)
    assert rel_err(rc1, rc2) < 1e-5 and rel_err(ra1, ra2) < 1e-5
    for x, y, name in zip(g1, g2, ("means", "quats", "scales", "opacities", "sh0", "shN")):
        assert rel_err(x, y) < 1e-4, name  # atomics: summation order differs between runs
