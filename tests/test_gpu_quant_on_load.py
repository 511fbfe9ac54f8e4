"""Quantise on load (compression_simulation fused into the attribute load of the projection / colour kernels) against
the unfused path of the same package: identical inputs, identical noise -- injected, or the same Philox counters.

What the unfused path is pinned to (reference ops.py / the reference CUDA kernels) is in test_gpu_parity.py; these tests
say that moving the fake quantisation into the consumers changes nothing: the forward is bit-identical, the gradients of
the raw parameters are the unfused ones (bit-identical where no atomics are involved).
"""
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
from gscodec_studio_b200 import _cabi  # noqa: E402
from gscodec_studio_b200._helper import load_test_data  # noqa: E402
from gscodec_studio_b200.compression_simulation import CompressionSimulation  # noqa: E402
from gscodec_studio_b200.compression_simulation.ops import QuantOnLoad, fake_quantize_ste  # noqa: E402
from gscodec_studio_b200.cuda._wrapper import (  # noqa: E402
    fully_fused_projection,
    fully_fused_projection_quant,
    view_colors_from_sh_split,
)

DEV = torch.device("cuda:0")
SH_C0 = 0.2820947917738781
BOUNDS = {"scales": (-10, 2), "quats": (-1, 1), "opacities": (-15, 15), "sh0": (-2, 4)}


def _raw_scene(C=2, seed=42, spread=1.0):
    means, quats, scales, opac, colors, viewmats, Ks, W, H = load_test_data(device=DEV, scene_grid=1, seed=seed)
    g = torch.Generator(device="cpu").manual_seed(5)
    raw = {
        "means": means,
        # a few percent of the entries sit OUTSIDE the quantiser's bounds, so the clamp gate is exercised
        "scales": torch.log(scales.clamp_min(1e-8)) * spread,
        "quats": quats * 1.3,
        "opacities": torch.logit(opac.clamp(1e-4, 1 - 1e-4)) * 2.5,
        "sh0": ((colors - 0.5) / SH_C0)[:, None, :] * 1.5,
        "shN": (torch.randn(means.shape[0], 15, 3, generator=g) * 0.05).to(DEV),
    }
    raw = {k: v.contiguous().requires_grad_(True) for k, v in raw.items()}
    return raw, viewmats[:C].contiguous(), Ks[:C].contiguous(), W, H


def _noise(raw, seed=11):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {k: (torch.rand(raw[k].shape, generator=g) - 0.5).to(DEV) for k in BOUNDS}


def _grads(params, loss):
    for p in params.values():
        p.grad = None
    loss.backward()
    return {k: p.grad.clone() for k, p in params.items() if p.grad is not None}


@pytest.mark.parametrize("injected", [True, False])
@pytest.mark.parametrize("camera_model", ["pinhole", "fisheye"])
def test_projection_quant_on_load_equals_unfused(injected, camera_model):
    raw, viewmats, Ks, W, H = _raw_scene(C=3)
    noise = _noise(raw) if injected else {k: None for k in BOUNDS}
    torch.manual_seed(123)
    lz = {k: QuantOnLoad(raw[k], *BOUNDS[k], 8, noise[k]) for k in ("quats", "scales", "opacities")}
    assert ((raw["scales"] < -10) | (raw["scales"] > 2)).any() and (raw["quats"].abs() > 1).any()
    kw = dict(eps2d=0.3, near_plane=0.01, far_plane=100.0, radius_clip=0.0, camera_model=camera_model)

    # unfused: the stand-alone kernels on the same draws, then the plain projection
    quats_q = lz["quats"].materialize()
    scales_a = torch.exp(lz["scales"]).materialize()
    opac_a = torch.sigmoid(lz["opacities"]).materialize()
    ref = fully_fused_projection(raw["means"], None, quats_q, scales_a, viewmats, Ks, W, H, calc_compensations=True, **kw)
    w = [torch.randn_like(t) for t in (ref[1], ref[2], ref[3], ref[4], opac_a)]
    g_ref = _grads(raw, sum((a * b).sum() for a, b in zip((ref[1], ref[2], ref[3], ref[4], opac_a), w)))

    n0 = _cabi.launch_count()
    out = fully_fused_projection_quant(raw["means"], lz["quats"], torch.exp(lz["scales"]), torch.sigmoid(lz["opacities"]),
                                       viewmats, Ks, W, H, calc_compensations=True, **kw)
    assert _cabi.launch_count() - n0 == 1, "the fused forward is ONE launch"
    names = ("radii", "means2d", "depths", "conics", "compensations")
    for name, a, b in zip(names, out[:5], ref):
        assert torch.equal(a, b), f"{name}: fused forward differs from the unfused one"
    assert torch.equal(out[5], opac_a), "activated opacities differ"
    g_new = _grads(raw, sum((a * b).sum() for a, b in zip((out[1], out[2], out[3], out[4], out[5]), w)))
    for k in ("means", "quats", "scales", "opacities"):
        assert torch.equal(g_new[k], g_ref[k]), (k, (g_new[k] - g_ref[k]).abs().max().item())
    # the clamp gate really bit: rows outside the bounds have zero gradient
    outside = (raw["scales"] < -10) | (raw["scales"] > 2)
    assert (g_new["scales"][outside] == 0).all()


@pytest.mark.parametrize("deg", [0, 3])
def test_sh_quant_on_load_equals_unfused(deg):
    raw, viewmats, Ks, W, H = _raw_scene(C=2)
    noise = _noise(raw)
    lz = QuantOnLoad(raw["sh0"], *BOUNDS["sh0"], 8, noise["sh0"])
    radii = fully_fused_projection(raw["means"].detach(), None, raw["quats"].detach(), torch.exp(raw["scales"].detach()),
                                   viewmats, Ks, W, H, far_plane=100.0)[0]
    campos = torch.linalg.inv(viewmats)[:, :3, 3].contiguous()
    ref = view_colors_from_sh_split(deg, raw["means"], campos, lz.materialize(), raw["shN"], radii)
    w = torch.randn_like(ref)
    g_ref = _grads(raw, (ref * w).sum())
    out = view_colors_from_sh_split(deg, raw["means"], campos, lz, raw["shN"], radii)
    assert torch.equal(out, ref)
    g_new = _grads(raw, (out * w).sum())
    for k in ("means", "sh0", "shN"):
        assert torch.equal(g_new[k], g_ref[k]), k
    outside = (raw["sh0"] < -2) | (raw["sh0"] > 4)
    assert outside.any() and (g_new["sh0"][outside] == 0).all()


@pytest.mark.parametrize("injected", [True, False])
def test_rasterization_with_lazy_attributes_equals_eager_step(injected):
    """The M2 step both ways: simulate_compression(fuse_into_render=True) + the reference's literal torch.exp /
    torch.sigmoid -> rasterization, against the eager path -- and no fake_quantize launch is left in the fused step."""
    raw, viewmats, Ks, W, H = _raw_scene(C=2)
    sim = CompressionSimulation(entropy_model_enable=False, device=DEV)
    noise = _noise(raw) if injected else None
    target = torch.rand(2, H, W, 3, device=DEV)

    def step(fused):
        torch.manual_seed(77)  # same Philox offsets both ways
        new, _ = sim.simulate_compression(raw, 100, noise=noise, fuse_into_render=fused)
        rc, ra, meta = gs.rasterization(new["means"], new["quats"], torch.exp(new["scales"]), torch.sigmoid(new["opacities"]),
                                        (new["sh0"], new["shN"]), viewmats, Ks, W, H, sh_degree=3, packed=False,
                                        near_plane=0.01, far_plane=100.0)
        return rc, ra, meta, _grads(raw, (rc - target).abs().mean() + 0.1 * ra.mean())

    rc0, ra0, meta0, g0 = step(False)
    _cabi.enable_stage_timing(True)
    rc1, ra1, meta1, g1 = step(True)
    calls = set(_cabi.stage_timing_ms())
    _cabi.enable_stage_timing(False)
    assert not any("fake_quantize" in c for c in calls), calls
    assert {"gsb200_projection_quant_fwd", "gsb200_projection_quant_bwd", "gsb200_sh_view_split_quant_fwd",
            "gsb200_sh_view_split_quant_bwd"} <= calls, calls
    assert torch.equal(meta1["radii"], meta0["radii"]) and torch.equal(meta1["conics"], meta0["conics"])
    assert torch.equal(meta1["isect_ids"], meta0["isect_ids"]) and torch.equal(meta1["flatten_ids"], meta0["flatten_ids"])
    assert torch.equal(rc1, rc0) and torch.equal(ra1, ra0), "fused render differs from the eager one"
    for k in raw:
        scale = g0[k].abs().max().item()
        err = (g1[k] - g0[k]).abs().max().item()
        assert err <= 2e-5 * scale, (k, err, scale)  # the blend backward accumulates with float atomics in both runs


def test_lazy_object_falls_back_to_the_materialised_tensor():
    raw, *_ = _raw_scene(C=1)
    noise = _noise(raw)
    lz = QuantOnLoad(raw["scales"], *BOUNDS["scales"], 8, noise["scales"])
    ref = fake_quantize_ste(raw["scales"], *BOUNDS["scales"], 8, noise=noise["scales"])["output_value"]
    assert torch.equal(lz.materialize(), ref)
    assert torch.equal(torch.exp(lz).materialize(), torch.exp(ref))
    assert torch.equal(lz + 1.0, ref + 1.0) and torch.equal(lz[3:7], ref[3:7]) and torch.equal(torch.cat([lz, lz], 1), torch.cat([ref, ref], 1))
    assert torch.equal(lz.mean(), ref.mean()) and lz.shape == ref.shape and len(lz) == len(ref)
    # every other rasterization mode materialises: packed here
    viewmats, Ks, W, H = _raw_scene(C=1)[1:]
    lq = QuantOnLoad(raw["quats"], *BOUNDS["quats"], 8, noise["quats"])
    lo = QuantOnLoad(raw["opacities"], *BOUNDS["opacities"], 8, noise["opacities"])
    colors = torch.rand(raw["means"].shape[0], 3, device=DEV)
    a = gs.rasterization(raw["means"], lq, torch.exp(lz), torch.sigmoid(lo), colors, viewmats, Ks, W, H, packed=True)[0]
    b = gs.rasterization(raw["means"], lq.materialize(), torch.exp(ref), torch.sigmoid(lo.materialize()), colors, viewmats, Ks,
                         W, H, packed=True)[0]
    assert torch.equal(a, b)
