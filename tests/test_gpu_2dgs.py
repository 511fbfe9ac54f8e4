"""2D Gaussian (surfel) family against the UNMODIFIED reference CUDA kernels compiled under oracle/_ref (live, same GPU,
same inputs), modelled on the reference's tests/test_2dgs.py:47-393 -- with tighter bounds than that file uses for its
CUDA-vs-torch comparison, because here both sides evaluate the same expressions:

    projection   radii EQUAL; means2d / depths / ray_transforms / normals equal to 1e-6 relative (bit-equal expected)
    blend        last_ids / median_ids EQUAL; colours / alphas / normals / distortion / median depth <= 1e-4 relative
    gradients    <= 1e-3 relative (max-normalised), the north-star tolerance
"""
import math
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
from gscodec_studio_b200._helper import load_test_data  # noqa: E402
from gscodec_studio_b200.cuda._wrapper import _isect_tiles_with_offsets  # noqa: E402
from gscodec_studio_b200.cuda._wrapper_2dgs import (  # noqa: E402
    accumulate_2dgs,
    fully_fused_projection_2dgs,
    rasterize_to_indices_in_range_2dgs,
    rasterize_to_pixels_2dgs,
)

DEV = torch.device("cuda:0")


def _ref():
    path = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(path, "gsplat_ref_cuda.so")):
        pytest.skip("oracle/_ref/gsplat_ref_cuda.so was not built (make -C oracle ref)")
    if path not in sys.path:
        sys.path.insert(0, path)
    import gsplat_ref_cuda as ref

    if not hasattr(ref, "fully_fused_projection_fwd_2dgs"):
        pytest.skip("oracle/_ref/gsplat_ref_cuda.so predates the 2DGS ops (make -C oracle ref)")
    return ref


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _scene(C=3, n=40_000, seed=42):
    means, quats, scales, opac, colors, viewmats, Ks, W, H = load_test_data(device=DEV, scene_grid=1, seed=seed)
    idx = torch.randperm(means.shape[0], generator=torch.Generator().manual_seed(1))[:n].to(DEV)
    means, quats, opac, colors = means[idx].contiguous(), quats[idx].contiguous(), opac[idx].contiguous(), colors[idx].contiguous()
    scales = (scales[idx] * 3.0).contiguous()  # surfels: a bit larger than the 3D splats of the fixture
    return means, quats, scales, opac, colors, viewmats[:C].contiguous(), Ks[:C].contiguous(), W, H


def _project(s, near=0.2, far=100.0, clip=0.0):
    means, quats, scales, opac, colors, viewmats, Ks, W, H = s
    return fully_fused_projection_2dgs(means, quats, scales, viewmats, Ks, W, H, 0.3, near, far, clip)


def test_projection_2dgs_matches_reference_cuda():
    ref = _ref()
    s = _scene()
    means, quats, scales, opac, colors, viewmats, Ks, W, H = s
    for t in (means, quats, scales):
        t.requires_grad_(True)
    radii, means2d, depths, rt, normals = _project(s, clip=2.0)
    r_radii, r_means2d, r_depths, r_rt, r_normals = ref.fully_fused_projection_fwd_2dgs(
        means.detach(), quats.detach(), scales.detach(), viewmats, Ks, W, H, 0.3, 0.2, 100.0, 2.0)
    # (entries the reference culls are left uninitialised by its kernel except for radii = 0; a degenerate transform with
    # distance == 0 is left entirely unwritten there -- none occurs in this scene)
    assert torch.equal(radii, r_radii), f"{(radii != r_radii).sum().item()} radii differ"
    valid = radii > 0
    assert valid.sum() > 5000
    for name, a, b in (("means2d", means2d, r_means2d), ("depths", depths, r_depths), ("ray_transforms", rt, r_rt),
                       ("normals", normals, r_normals)):
        assert rel(a[valid], b[valid]) <= 1e-6, (name, rel(a[valid], b[valid]))
        print(name, "bit-equal fraction", (a[valid] == b[valid]).float().mean().item())
    assert (means2d[~valid] == 0).all() and (rt[~valid] == 0).all() and (normals[~valid] == 0).all()

    g = torch.Generator(device="cpu").manual_seed(3)
    v = [torch.randn(t.shape, generator=g).to(DEV) * valid.reshape(valid.shape + (1,) * (t.dim() - 2))
         for t in (means2d, depths, normals, rt)]
    (means2d * v[0]).sum().add((depths * v[1]).sum()).add((normals * v[2]).sum()).add((rt * v[3]).sum()).backward()
    r_v_means, r_v_quats, r_v_scales, _ = ref.fully_fused_projection_bwd_2dgs(
        means.detach(), quats.detach(), scales.detach(), viewmats, Ks, W, H, r_radii, r_rt, v[0].contiguous(), v[1].contiguous(),
        v[2].contiguous(), v[3].contiguous(), False)
    assert rel(means.grad, r_v_means) <= 1e-4, rel(means.grad, r_v_means)
    assert rel(quats.grad, r_v_quats) <= 1e-4, rel(quats.grad, r_v_quats)
    assert rel(scales.grad, r_v_scales) <= 1e-4, rel(scales.grad, r_v_scales)
    assert (scales.grad[:, 2] == 0).all()


def test_projection_2dgs_packed_equals_dense():
    s = _scene(C=2)
    means, quats, scales, opac, colors, viewmats, Ks, W, H = s
    for t in (means, quats, scales):
        t.requires_grad_(True)
    radii, means2d, depths, rt, normals = _project(s)
    cam, gid, p_radii, p_means2d, p_depths, p_rt, p_normals = fully_fused_projection_2dgs(
        means, quats, scales, viewmats, Ks, W, H, 0.3, 0.2, 100.0, 0.0, packed=True)
    sel = radii > 0
    assert torch.equal(torch.stack(torch.nonzero(sel, as_tuple=True)), torch.stack([cam, gid]))
    assert torch.equal(p_radii, radii[sel]) and torch.equal(p_means2d, means2d[sel]) and torch.equal(p_rt, rt[sel])
    w = [torch.randn_like(t) for t in (p_means2d, p_depths, p_normals, p_rt)]
    loss_p = sum((a * b).sum() for a, b in zip((p_means2d, p_depths, p_normals, p_rt), w))
    gp = torch.autograd.grad(loss_p, (means, quats, scales))
    loss_d = sum((a[sel] * b).sum() for a, b in zip((means2d, depths, normals, rt), w))
    gd = torch.autograd.grad(loss_d, (means, quats, scales))
    for a, b in zip(gp, gd):
        assert torch.equal(a, b)


def _blend_inputs(channels=3, C=2, tile=16):
    s = _scene(C=C)
    means, quats, scales, opac, colors, viewmats, Ks, W, H = s
    radii, means2d, depths, rt, normals = [t.detach() for t in _project(s)]
    tw, th = math.ceil(W / tile), math.ceil(H / tile)
    _, isect_ids, flatten_ids, offsets = _isect_tiles_with_offsets(means2d, radii, depths, tile, tw, th)
    g = torch.Generator(device="cpu").manual_seed(9)
    feats = torch.rand(C, means.shape[0], channels, generator=g).to(DEV)
    feats[..., -1] = depths  # the last channel is what the distortion / median outputs read as depth
    opacities = opac[None].repeat(C, 1).contiguous()
    return means2d, rt, feats, opacities, normals, offsets, flatten_ids, W, H


@pytest.mark.parametrize("channels,bg,absgrad,C", [(3, False, False, 1), (4, True, True, 1), (1, False, False, 1), (8, True, False, 1),
                                                   (4, True, False, 2)])
def test_rasterize_to_pixels_2dgs_matches_reference_cuda(channels, bg, absgrad, C):
    """C = 2 runs without a distortion gradient: the reference's backward kernel reads `render_colors` (the rendered depth the
    distortion term needs) without the camera offset it applies to every other per-pixel tensor
    (rasterize_to_pixels_2dgs_bwd.cu:79-97 vs :203), i.e. camera 0's depth for every camera; ours reads the right camera."""
    ref = _ref()
    tile = 16
    means2d, rt, feats, opacities, normals, offsets, flatten_ids, W, H = _blend_inputs(channels, C=C)
    C = means2d.shape[0]
    backgrounds = torch.rand(C, channels, device=DEV) if bg else None
    leaves = [t.clone().requires_grad_(True) for t in (means2d, rt, feats, opacities, normals)]
    densify = torch.zeros_like(means2d, requires_grad=True)
    out = rasterize_to_pixels_2dgs(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], densify, W, H, tile, offsets,
                                   flatten_ids, backgrounds=backgrounds, absgrad=absgrad)
    r = ref.rasterize_to_pixels_fwd_2dgs(means2d, rt, feats, opacities, normals, backgrounds, None, W, H, tile, offsets, flatten_ids)
    r_colors, r_alphas, r_normals, r_distort, r_median, r_last, r_medid = r
    names = ("colors", "alphas", "normals", "distort", "median")
    for name, a, b in zip(names, out, (r_colors, r_alphas, r_normals, r_distort, r_median)):
        assert rel(a, b) <= 1e-4, (name, rel(a, b))
        print(name, "max rel err", rel(a, b))
    assert (out[1] > 0.5).float().mean() > 0.2, "the scene should cover a good part of the image"

    g = torch.Generator(device="cpu").manual_seed(4)
    v = [torch.randn(t.shape, generator=g).to(DEV) for t in out]
    if C > 1:
        v[3] = torch.zeros_like(v[3])
    sum((a * b).sum() for a, b in zip(out, v)).backward()
    rb = ref.rasterize_to_pixels_bwd_2dgs(means2d, rt, feats, opacities, normals, torch.zeros_like(means2d), backgrounds, None, W,
                                          H, tile, offsets, flatten_ids, r_colors, r_alphas, r_last, r_medid, v[0].contiguous(),
                                          v[1].contiguous(), v[2].contiguous(), v[3].contiguous(), v[4].contiguous(), absgrad)
    r_abs, r_v_means2d, r_v_rt, r_v_colors, r_v_opac, r_v_normals, r_v_densify = rb
    errs = {}
    for name, a, b in (("v_means2d", leaves[0].grad, r_v_means2d), ("v_ray_transforms", leaves[1].grad, r_v_rt),
                       ("v_colors", leaves[2].grad, r_v_colors), ("v_opacities", leaves[3].grad, r_v_opac),
                       ("v_normals", leaves[4].grad, r_v_normals)):
        errs[name] = rel(a, b)
        worst = (a - b).abs().flatten().argmax().item()
        print(name, "max rel err", errs[name], "| worst entry: ours", a.flatten()[worst].item(), "ref", b.flatten()[worst].item(),
              "| max |ref|", b.abs().max().item(), "| nonzero ours / ref", (a != 0).sum().item(), (b != 0).sum().item())
    assert all(e <= 1e-3 for e in errs.values()), errs
    if absgrad:
        assert rel(leaves[0].absgrad, r_abs) <= 1e-3
    # densification signal: the deterministic value the reference's racy read converges to
    want = torch.stack([leaves[1].grad[..., 0, 2], leaves[1].grad[..., 1, 2]], -1) * rt[..., 2, 2:3]
    assert torch.allclose(densify.grad, want, rtol=1e-6, atol=0)
    touched = (r_v_densify != 0).any(-1)
    assert ((densify.grad != 0).any(-1) | ~touched).all()


def test_rasterize_to_pixels_2dgs_masks_and_small_tiles():
    tile = 8
    s = _scene(C=1, n=20_000)
    means, quats, scales, opac, colors, viewmats, Ks, W, H = s
    radii, means2d, depths, rt, normals = [t.detach() for t in _project(s)]
    tw, th = math.ceil(W / tile), math.ceil(H / tile)
    _, _, flatten_ids, offsets = _isect_tiles_with_offsets(means2d, radii, depths, tile, tw, th)
    feats = torch.cat([colors[None], depths[..., None]], -1).contiguous()
    opacities = opac[None].contiguous()
    masks = torch.rand(1, th, tw, device=DEV) > 0.3
    bgs = torch.rand(1, 4, device=DEV)
    full = rasterize_to_pixels_2dgs(means2d, rt, feats, opacities, normals, torch.zeros_like(means2d), W, H, tile, offsets,
                                    flatten_ids, backgrounds=bgs)
    part = rasterize_to_pixels_2dgs(means2d, rt, feats, opacities, normals, torch.zeros_like(means2d), W, H, tile, offsets,
                                    flatten_ids, backgrounds=bgs, masks=masks)
    pix = masks.repeat_interleave(tile, 1).repeat_interleave(tile, 2)[:, :H, :W]
    for a, b in zip(full, part):
        assert torch.equal(a[pix], b[pix])
    assert torch.equal(part[0][~pix], bgs[0].expand_as(part[0][~pix])) and (part[1][~pix] == 0).all()


def test_indices_in_range_2dgs_matches_reference_cuda_and_accumulate():
    ref = _ref()
    tile = 16
    means2d, rt, feats, opacities, normals, offsets, flatten_ids, W, H = _blend_inputs(3, C=2)
    C = means2d.shape[0]
    trans = torch.ones(C, H, W, device=DEV)
    ours = rasterize_to_indices_in_range_2dgs(0, 10_000, trans, means2d, rt, opacities, W, H, tile, offsets, flatten_ids)
    r_gids, r_idx = ref.rasterize_to_indices_in_range_2dgs(0, 10_000, trans, means2d, rt, opacities, W, H, tile, offsets, flatten_ids)
    assert torch.equal(ours[0], r_gids)
    assert torch.equal(ours[1], r_idx % (W * H)) and torch.equal(ours[2], r_idx // (W * H))
    # a slice in the middle, starting from a non-trivial transmittance
    trans2 = torch.rand(C, H, W, device=DEV) * 0.9 + 0.1
    ours2 = rasterize_to_indices_in_range_2dgs(1, 3, trans2, means2d, rt, opacities, W, H, tile, offsets, flatten_ids)
    r2 = ref.rasterize_to_indices_in_range_2dgs(1, 3, trans2, means2d, rt, opacities, W, H, tile, offsets, flatten_ids)
    assert torch.equal(ours2[0], r2[0]) and torch.equal(ours2[1] + ours2[2] * W * H, r2[1])
    # the torch rasterizer built on those lists reproduces the fused blend (reference: _torch_impl_2dgs.py:171-265)
    colors, alphas, nrm = accumulate_2dgs(means2d, rt, opacities, feats, normals, ours[0], ours[1], ours[2], W, H)
    fused = rasterize_to_pixels_2dgs(means2d, rt, feats, opacities, normals, torch.zeros_like(means2d), W, H, tile, offsets,
                                     flatten_ids)
    # (torch.exp vs the kernels' __expf, and float32 products of up to a few hundred (1 - alpha) factors)
    assert rel(colors, fused[0]) <= 1e-3 and rel(alphas, fused[1]) <= 1e-3 and rel(nrm, fused[2]) <= 1e-3


def test_rasterize_to_pixels_2dgs_gradients_match_torch_autograd():
    """Independent check of the blend backward: autograd through the torch rasterizer (accumulate_2dgs on the contributor
    lists, float64) -- the reference's own way of testing its kernels (tests/test_2dgs.py:234-393)."""
    tile = 16
    means2d, rt, feats, opacities, normals, offsets, flatten_ids, W, H = _blend_inputs(3, C=1)
    trans = torch.ones(1, H, W, device=DEV)
    gids, pids, cids = rasterize_to_indices_in_range_2dgs(0, 10_000, trans, means2d, rt, opacities, W, H, tile, offsets, flatten_ids)
    leaves = [t.clone().requires_grad_(True) for t in (means2d, rt, feats, opacities, normals)]
    out = rasterize_to_pixels_2dgs(*leaves, torch.zeros_like(means2d), W, H, tile, offsets, flatten_ids)
    g = torch.Generator(device="cpu").manual_seed(4)
    v = [torch.randn(t.shape, generator=g).to(DEV) for t in out[:3]]
    sum((a * b).sum() for a, b in zip(out[:3], v)).backward()
    leaves64 = [t.double().clone().requires_grad_(True) for t in (means2d, rt, feats, opacities, normals)]
    out64 = accumulate_2dgs(leaves64[0], leaves64[1], leaves64[3], leaves64[2], leaves64[4], gids, pids, cids, W, H)
    sum((a * b.double()).sum() for a, b in zip(out64, v)).backward()
    errs = {}
    for name, a, b in zip(("v_means2d", "v_ray_transforms", "v_colors", "v_opacities", "v_normals"), leaves, leaves64):
        errs[name] = rel(a.grad, b.grad)
        print(name, "vs torch autograd (float64): max rel err", errs[name])
    assert all(e <= 1e-3 for e in errs.values()), errs


@pytest.mark.parametrize("packed", [False, True])
def test_rasterization_2dgs_end_to_end(packed):
    means, quats, scales, opac, colors, viewmats, Ks, W, H = _scene(C=2, n=30_000)
    params = [t.clone().requires_grad_(True) for t in (means, quats, scales, opac, colors)]
    out = gs.rasterization_2dgs(*params, viewmats, Ks, W, H, near_plane=0.2, far_plane=100.0, render_mode="RGB+ED", packed=packed,
                                distloss=True, absgrad=True)
    rc, ra, rn, rnd, dist, med, meta = out
    assert rc.shape == (2, H, W, 4) and ra.shape == (2, H, W, 1) and rn.shape == (2, H, W, 3) and rnd.shape == (2, H, W, 3)
    assert dist.shape == (2, H, W, 1) and med.shape == (2, H, W, 1)
    for k in ("camera_ids", "gaussian_ids", "radii", "means2d", "depths", "ray_transforms", "opacities", "normals", "tile_width",
              "tile_height", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets", "width", "height", "tile_size",
              "n_cameras", "render_distort", "gradient_2dgs"):
        assert k in meta, k
    meta["means2d"].retain_grad()
    loss = rc[..., :3].mean() + 0.1 * (1 - (rn * rnd).sum(-1)).mean() + 0.01 * dist.mean() + 0.01 * med.mean() + ra.mean()
    loss.backward()
    for p in params:
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert meta["gradient_2dgs"].grad is not None and meta["gradient_2dgs"].grad.abs().sum() > 0
    assert hasattr(meta["means2d"], "absgrad")
    if packed:
        dense = gs.rasterization_2dgs(means, quats, scales, opac, colors, viewmats, Ks, W, H, near_plane=0.2, far_plane=100.0,
                                      render_mode="RGB+ED", packed=False)
        assert torch.equal(dense[0], rc.detach()) and torch.equal(dense[1], ra.detach())
    with pytest.raises(NotImplementedError):
        gs.rasterization_2dgs_inria_wrapper()
