"""The single-GPU colour stage runs on a side stream (rendering._colour_stream): forward beside the tile intersection,
backward -- autograd runs a node's backward on its forward's stream -- beside the projection backward.  Moving it must
change nothing: images and colour-path gradients bit-identical to the same call with GSB200_COLOR_STREAM=0, geometry
gradients equal up to the order of the blend's float atomics; and the join must hold when the caller overwrites the
parameters right after the call (the optimizer step of a training loop)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)
sys.path.insert(0, ROOT)

import gscodec_studio_b200 as gs  # noqa: E402
from gscodec_studio_b200 import rendering  # noqa: E402
from gscodec_studio_b200._helper import load_test_data  # noqa: E402

DEV = torch.device("cuda:0")
SH_C0 = 0.2820947917738781


def _scene(C=2):
    means, quats, scales, opac, colors, viewmats, Ks, W, H = load_test_data(device=DEV, scene_grid=1)
    g = torch.Generator(device="cpu").manual_seed(3)
    sh0 = ((colors - 0.5) / SH_C0)[:, None, :].contiguous()
    shN = (torch.randn(means.shape[0], 15, 3, generator=g) * 0.05).to(DEV)
    params = {"means": means, "quats": quats, "scales": scales, "opacities": opac, "sh0": sh0, "shN": shN}
    return {k: v.clone().requires_grad_(True) for k, v in params.items()}, viewmats[:C].contiguous(), Ks[:C].contiguous(), W, H


def _step(params, viewmats, Ks, W, H, pair, weight, sync=True):
    for p in params.values():
        p.grad = None
    colors = (params["sh0"], params["shN"]) if pair else torch.cat([params["sh0"], params["shN"]], dim=1)
    img, alpha, meta = gs.rasterization(params["means"], params["quats"], params["scales"], params["opacities"], colors,
                                        viewmats, Ks, W, H, sh_degree=3, packed=False)
    ((img * weight).sum() + alpha.sum() * 0.1).backward()
    if sync:
        torch.cuda.synchronize()
    return img.detach().clone(), alpha.detach().clone(), {k: p.grad.detach().clone() for k, p in params.items()}


@pytest.mark.parametrize("pair", [True, False])
def test_side_stream_changes_nothing(pair, monkeypatch):
    params, viewmats, Ks, W, H = _scene()
    weight = torch.rand(viewmats.shape[0], H, W, 3, device=DEV)
    monkeypatch.setenv("GSB200_COLOR_STREAM", "0")
    assert rendering._colour_stream(DEV) is None
    img0, a0, g0 = _step(params, viewmats, Ks, W, H, pair, weight)
    monkeypatch.setenv("GSB200_COLOR_STREAM", "1")
    assert rendering._colour_stream(DEV) is not None
    for _ in range(3):  # more than once: the allocator hands the side stream's blocks out again
        img1, a1, g1 = _step(params, viewmats, Ks, W, H, pair, weight)
        assert torch.equal(img0, img1) and torch.equal(a0, a1)
        for k in g0:
            scale = g0[k].abs().max().item()
            err = (g0[k] - g1[k]).abs().max().item() / max(scale, 1e-20)
            assert err < 2e-5, f"{k}: {err:.2e}"  # float atomics of the blend backward only


def test_join_holds_against_an_immediate_parameter_update():
    """Training-loop order: backward, then the parameters are overwritten on the main stream at once.  The colour kernels
    of the NEXT step fork after that update; the ones of THIS step must have read the old values."""
    params, viewmats, Ks, W, H = _scene(C=1)
    weight = torch.rand(1, H, W, 3, device=DEV)
    ref_img, _, ref_g = _step(params, viewmats, Ks, W, H, True, weight)
    saved = {k: p.detach().clone() for k, p in params.items()}
    for _ in range(4):
        img, _, g = _step(params, viewmats, Ks, W, H, True, weight, sync=False)  # no host wait before the update
        with torch.no_grad():
            for k in ("sh0", "shN", "means"):  # clobber what the side stream reads, then restore: both on the main stream
                params[k].mul_(0.0)
            for k in ("sh0", "shN", "means"):
                params[k].copy_(saved[k])
        assert torch.equal(img, ref_img)
        for k in ("sh0", "shN", "means"):  # (the colour gradients come out of the blend's float atomics: equal up to their order)
            err = (g[k] - ref_g[k]).abs().max().item() / max(ref_g[k].abs().max().item(), 1e-20)
            assert err < 2e-5, f"{k}: {err:.2e}"
