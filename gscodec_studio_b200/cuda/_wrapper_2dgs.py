"""2D Gaussian (surfel) operators over the libgsb200 C-ABI (csrc/surfel.cu).

Public names, argument order, defaults, asserts and returns are those of the reference's ``gsplat/cuda/_wrapper.py``:

    fully_fused_projection_2dgs          _wrapper.py:1260-1361, _FullyFusedProjection2DGS :1364-1468,
                                                                _FullyFusedProjectionPacked2DGS :1471-1623
    rasterize_to_pixels_2dgs             _wrapper.py:1626-1749, _RasterizeToPixels2DGS :1828-1992
    rasterize_to_indices_in_range_2dgs   _wrapper.py:1752-1825

Differences (documented, behaviour-preserving):
  * entries the reference leaves uninitialised (culled (camera, Gaussian) pairs, masked tiles) are zero;
  * `packed=True` projects densely and compacts (memory C * N inside the op, [nnz, .] outside, same row order);
  * `v_densify` is computed from the finished gradient sums (the reference reads them while other warps still accumulate);
  * the depth-distortion gradient reads the rendered depth of ITS camera (the reference's backward kernel forgets the camera
    offset of `render_colors`, rasterize_to_pixels_2dgs_bwd.cu:79-97 vs :203, so with C > 1 every camera uses camera 0's);
  * channel counts are padded to 1, 2, 3, 4, 8, 16 or 32 with ZEROS (the reference pads with uninitialised memory, which the
    depth-distortion term then reads as "depth"); more than 32 channels raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .._cabi import call, device_guard, ptr, stream_ptr
from ._wrapper import _f32

_CHANNELS_2DGS = (1, 2, 3, 4, 8, 16, 32)


def fully_fused_projection_2dgs(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]
    scales: Tensor,  # [N, 3]
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    packed: bool = False,
    sparse_grad: bool = False,
):
    """Ray-splat transforms, bounding boxes and 2D means of 2D Gaussians.

    dense:  (radii i32 [C,N], means2d [C,N,2], depths [C,N], ray_transforms [C,N,3,3], normals [C,N,3])
    packed: (camera_ids [nnz], gaussian_ids [nnz], radii [nnz], means2d [nnz,2], depths [nnz], ray_transforms [nnz,3,3],
             normals [nnz,3])"""
    C = viewmats.size(0)
    N = means.size(0)
    assert means.size() == (N, 3), means.size()
    assert viewmats.size() == (C, 4, 4), viewmats.size()
    assert Ks.size() == (C, 3, 3), Ks.size()
    assert quats is not None, "quats is required"
    assert scales is not None, "scales is required"
    assert quats.size() == (N, 4), quats.size()
    assert scales.size() == (N, 3), scales.size()
    if sparse_grad:
        assert packed, "sparse_grad is only supported when packed is True"
    args = (means.contiguous(), quats.contiguous(), scales.contiguous(), viewmats.contiguous(), Ks.contiguous(), width, height)
    if packed:
        return _FullyFusedProjectionPacked2DGS.apply(*args, near_plane, far_plane, radius_clip, sparse_grad)
    return _FullyFusedProjection2DGS.apply(*args, eps2d, near_plane, far_plane, radius_clip)


def _projection_2dgs_fwd(means, quats, scales, viewmats, Ks, width, height, near_plane, far_plane, radius_clip):
    C, N = viewmats.shape[0], means.shape[0]
    dev = means.device
    radii = torch.empty((C, N), dtype=torch.int32, device=dev)
    means2d = torch.empty((C, N, 2), dtype=torch.float32, device=dev)
    depths = torch.empty((C, N), dtype=torch.float32, device=dev)
    ray_transforms = torch.empty((C, N, 3, 3), dtype=torch.float32, device=dev)
    normals = torch.empty((C, N, 3), dtype=torch.float32, device=dev)
    with device_guard(dev):
        call("gsb200_projection_2dgs_fwd", C, N, ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), width, height,
             near_plane, far_plane, radius_clip, ptr(radii), ptr(means2d), ptr(depths), ptr(ray_transforms), ptr(normals),
             stream_ptr(dev))
    return radii, means2d, depths, ray_transforms, normals


def _projection_2dgs_bwd(means, quats, scales, viewmats, Ks, radii, ray_transforms, v_means2d, v_depths, v_normals,
                         v_ray_transforms):
    C, N = viewmats.shape[0], means.shape[0]
    dev = means.device
    v_means, v_quats, v_scales = torch.empty_like(means), torch.empty_like(quats), torch.empty_like(scales)
    with device_guard(dev):
        call("gsb200_projection_2dgs_bwd", C, N, ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), ptr(radii),
             ptr(ray_transforms), ptr(_f32(v_means2d.contiguous())), ptr(_f32(v_depths.contiguous())),
             ptr(_f32(v_normals.contiguous())), ptr(_f32(v_ray_transforms.contiguous())), ptr(v_means), ptr(v_quats),
             ptr(v_scales), stream_ptr(dev))
    return v_means, v_quats, v_scales


class _FullyFusedProjection2DGS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip):
        means, quats, scales, viewmats, Ks = map(_f32, (means, quats, scales, viewmats, Ks))
        radii, means2d, depths, ray_transforms, normals = _projection_2dgs_fwd(
            means, quats, scales, viewmats, Ks, width, height, near_plane, far_plane, radius_clip)
        ctx.save_for_backward(means, quats, scales, viewmats, Ks, radii, ray_transforms)
        ctx.mark_non_differentiable(radii)
        return radii, means2d, depths, ray_transforms, normals

    @staticmethod
    def backward(ctx, v_radii, v_means2d, v_depths, v_ray_transforms, v_normals):
        means, quats, scales, viewmats, Ks, radii, ray_transforms = ctx.saved_tensors
        if ctx.needs_input_grad[3]:
            raise RuntimeError("fully_fused_projection_2dgs: camera gradients are not implemented (nor are they in the "
                               "reference: fully_fused_projection_2dgs_bwd.cu never writes v_viewmats)")
        v_means, v_quats, v_scales = _projection_2dgs_bwd(means, quats, scales, viewmats, Ks, radii, ray_transforms, v_means2d,
                                                          v_depths, v_normals, v_ray_transforms)
        return (v_means if ctx.needs_input_grad[0] else None, v_quats if ctx.needs_input_grad[1] else None,
                v_scales if ctx.needs_input_grad[2] else None, None, None, None, None, None, None, None, None)


class _FullyFusedProjectionPacked2DGS(torch.autograd.Function):
    """Dense projection + compaction to the rows with radii > 0, in (camera, Gaussian) order (the reference's row order,
    fully_fused_projection_packed_2dgs_fwd.cu); the backward scatters the packed gradients back and runs the dense kernel."""

    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, near_plane, far_plane, radius_clip, sparse_grad):
        means, quats, scales, viewmats, Ks = map(_f32, (means, quats, scales, viewmats, Ks))
        radii, means2d, depths, ray_transforms, normals = _projection_2dgs_fwd(
            means, quats, scales, viewmats, Ks, width, height, near_plane, far_plane, radius_clip)
        sel = radii > 0
        camera_ids, gaussian_ids = torch.nonzero(sel, as_tuple=True)
        ctx.save_for_backward(means, quats, scales, viewmats, Ks, radii, ray_transforms, camera_ids, gaussian_ids)
        ctx.sparse_grad = sparse_grad
        out = (camera_ids, gaussian_ids, radii[sel], means2d[sel], depths[sel], ray_transforms[sel], normals[sel])
        ctx.mark_non_differentiable(*out[:3])
        return out

    @staticmethod
    def backward(ctx, v_camera_ids, v_gaussian_ids, v_radii, v_means2d, v_depths, v_ray_transforms, v_normals):
        means, quats, scales, viewmats, Ks, radii, ray_transforms, camera_ids, gaussian_ids = ctx.saved_tensors
        if ctx.needs_input_grad[3]:
            raise RuntimeError("fully_fused_projection_2dgs: camera gradients are not implemented")
        C, N = radii.shape

        def dense(v, tail):
            out = torch.zeros((C, N) + tail, dtype=torch.float32, device=radii.device)
            out[camera_ids, gaussian_ids] = v.to(torch.float32)
            return out

        v_means, v_quats, v_scales = _projection_2dgs_bwd(
            means, quats, scales, viewmats, Ks, radii, ray_transforms, dense(v_means2d, (2,)), dense(v_depths, ()),
            dense(v_normals, (3,)), dense(v_ray_transforms, (3, 3)))

        def out(t, need):
            if not need:
                return None
            if ctx.sparse_grad:  # same values as the reference's COO gradient once densified (_wrapper.py:1580-1608)
                rows = torch.unique(gaussian_ids)
                return torch.sparse_coo_tensor(rows[None], t[rows], size=t.shape, is_coalesced=True)
            return t

        return (out(v_means, ctx.needs_input_grad[0]), out(v_quats, ctx.needs_input_grad[1]),
                out(v_scales, ctx.needs_input_grad[2]), None, None, None, None, None, None, None, None)


def rasterize_to_pixels_2dgs(
    means2d: Tensor,  # [C, N, 2] or [nnz, 2]
    ray_transforms: Tensor,  # [C, N, 3, 3] or [nnz, 3, 3]
    colors: Tensor,  # [C, N, channels] or [nnz, channels]
    opacities: Tensor,  # [C, N] or [nnz]
    normals: Tensor,  # [C, N, 3] or [nnz, 3]
    densify: Tensor,  # [C, N, 2] or [nnz, 2]
    image_width: int,
    image_height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C, tile_height, tile_width]
    flatten_ids: Tensor,  # [n_isects]
    backgrounds: Optional[Tensor] = None,  # [C, channels]
    masks: Optional[Tensor] = None,  # [C, tile_height, tile_width]
    packed: bool = False,
    absgrad: bool = False,
    distloss: bool = False,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Rasterizes 2D Gaussians to pixels.  -> (render_colors [C,H,W,channels], render_alphas [C,H,W,1], render_normals
    [C,H,W,3] (camera space), render_distort [C,H,W,1], render_median [C,H,W,1]).  `densify` is a dummy input whose
    gradient carries the densification signal of 2DGS."""
    C = isect_offsets.size(0)
    device = means2d.device
    if packed:
        nnz = means2d.size(0)
        assert means2d.shape == (nnz, 2), means2d.shape
        assert ray_transforms.shape == (nnz, 3, 3), ray_transforms.shape
        assert colors.shape[0] == nnz, colors.shape
        assert opacities.shape == (nnz,), opacities.shape
    else:
        N = means2d.size(1)
        assert means2d.shape == (C, N, 2), means2d.shape
        assert ray_transforms.shape == (C, N, 3, 3), ray_transforms.shape
        assert colors.shape[:2] == (C, N), colors.shape
        assert opacities.shape == (C, N), opacities.shape
    if backgrounds is not None:
        assert backgrounds.shape == (C, colors.shape[-1]), backgrounds.shape
        backgrounds = backgrounds.contiguous()
    channels = colors.shape[-1]
    if channels > 512 or channels == 0:
        raise ValueError(f"Unsupported number of color channels: {channels}")
    if channels > _CHANNELS_2DGS[-1]:
        raise ValueError(f"rasterize_to_pixels_2dgs: at most {_CHANNELS_2DGS[-1]} channels per call in this build, got {channels}")
    padded_channels = 0
    if channels not in _CHANNELS_2DGS:
        padded_channels = (1 << (channels - 1).bit_length()) - channels
        colors = torch.cat([colors, torch.zeros(*colors.shape[:-1], padded_channels, device=device)], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(*backgrounds.shape[:-1], padded_channels, device=device)], dim=-1)
    tile_height, tile_width = isect_offsets.shape[1:3]
    assert tile_height * tile_size >= image_height, f"Assert Failed: {tile_height} * {tile_size} >= {image_height}"
    assert tile_width * tile_size >= image_width, f"Assert Failed: {tile_width} * {tile_size} >= {image_width}"
    render_colors, render_alphas, render_normals, render_distort, render_median = _RasterizeToPixels2DGS.apply(
        means2d.contiguous(), ray_transforms.contiguous(), colors.contiguous(), opacities.contiguous(), normals.contiguous(),
        densify.contiguous(), backgrounds, masks, image_width, image_height, tile_size, isect_offsets.contiguous(),
        flatten_ids.contiguous(), absgrad, distloss)
    if padded_channels > 0:
        render_colors = render_colors[..., :-padded_channels]
    return render_colors, render_alphas, render_normals, render_distort, render_median


class _RasterizeToPixels2DGS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, ray_transforms, colors, opacities, normals, densify, backgrounds, masks, width, height, tile_size,
                isect_offsets, flatten_ids, absgrad, distloss):
        means2d_in = means2d
        means2d, ray_transforms, colors, opacities, normals = map(_f32, (means2d, ray_transforms, colors, opacities, normals))
        backgrounds = _f32(backgrounds)
        masks_u8 = None if masks is None else masks.to(torch.uint8).contiguous()
        C, th, tw = isect_offsets.shape
        D = colors.shape[-1]
        M = opacities.numel()
        dev = means2d.device
        new = torch.zeros if masks is not None else torch.empty
        render_colors = new((C, height, width, D), dtype=torch.float32, device=dev)
        render_alphas = new((C, height, width, 1), dtype=torch.float32, device=dev)
        render_normals = new((C, height, width, 3), dtype=torch.float32, device=dev)
        render_distort = new((C, height, width, 1), dtype=torch.float32, device=dev)
        render_median = new((C, height, width, 1), dtype=torch.float32, device=dev)
        last_ids = new((C, height, width), dtype=torch.int32, device=dev)
        median_ids = new((C, height, width), dtype=torch.int32, device=dev)
        with device_guard(dev):
            call("gsb200_rasterize_2dgs_fwd", C, M, flatten_ids.numel(), D, ptr(means2d), ptr(ray_transforms), ptr(colors),
                 ptr(opacities), ptr(normals), ptr(backgrounds), ptr(masks_u8), width, height, tile_size, tw, th,
                 ptr(isect_offsets), ptr(flatten_ids), ptr(render_colors), ptr(render_alphas), ptr(render_normals),
                 ptr(render_distort), ptr(render_median), ptr(last_ids), ptr(median_ids), stream_ptr(dev))
        ctx.save_for_backward(means2d, ray_transforms, colors, opacities, normals, densify, backgrounds, masks_u8, isect_offsets,
                              flatten_ids, render_colors, render_alphas, last_ids, median_ids)
        ctx.dims = (width, height, tile_size, absgrad)
        ctx.means2d_in = means2d_in if absgrad else None
        return render_colors, render_alphas, render_normals, render_distort, render_median

    @staticmethod
    def backward(ctx, v_render_colors, v_render_alphas, v_render_normals, v_render_distort, v_render_median):
        (means2d, ray_transforms, colors, opacities, normals, densify, backgrounds, masks_u8, isect_offsets, flatten_ids,
         render_colors, render_alphas, last_ids, median_ids) = ctx.saved_tensors
        width, height, tile_size, absgrad = ctx.dims
        C, th, tw = isect_offsets.shape
        D = colors.shape[-1]
        M = opacities.numel()
        dev = means2d.device
        grads = [_f32(v.contiguous()) for v in (v_render_colors, v_render_alphas, v_render_normals, v_render_distort,
                                               v_render_median)]
        v_means2d, v_ray_transforms = torch.empty_like(means2d), torch.empty_like(ray_transforms)
        v_colors, v_opacities, v_normals = torch.empty_like(colors), torch.empty_like(opacities), torch.empty_like(normals)
        v_densify = torch.zeros(densify.shape, dtype=torch.float32, device=dev)
        v_abs = torch.empty_like(means2d) if absgrad else None
        with device_guard(dev):
            call("gsb200_rasterize_2dgs_bwd", C, M, flatten_ids.numel(), D, ptr(means2d), ptr(ray_transforms), ptr(colors),
                 ptr(opacities), ptr(normals), ptr(backgrounds), ptr(masks_u8), width, height, tile_size, tw, th,
                 ptr(isect_offsets), ptr(flatten_ids), ptr(render_colors), ptr(render_alphas), ptr(last_ids), ptr(median_ids),
                 ptr(grads[0]), ptr(grads[1]), ptr(grads[2]), ptr(grads[3]), ptr(grads[4]), ptr(v_abs), ptr(v_means2d),
                 ptr(v_ray_transforms), ptr(v_colors), ptr(v_opacities), ptr(v_normals), ptr(v_densify), stream_ptr(dev))
        if absgrad:
            ctx.means2d_in.absgrad = v_abs
        v_backgrounds = None
        if ctx.needs_input_grad[6]:
            v_backgrounds = (grads[0] * (1.0 - render_alphas).float()).sum(dim=(1, 2))
        return (v_means2d, v_ray_transforms, v_colors, v_opacities, v_normals, v_densify, v_backgrounds, None, None, None, None,
                None, None, None, None)


@torch.no_grad()
def rasterize_to_indices_in_range_2dgs(
    range_start: int,
    range_end: int,
    transmittances: Tensor,  # [C, image_height, image_width]
    means2d: Tensor,  # [C, N, 2]
    ray_transforms: Tensor,  # [C, N, 3, 3]
    opacities: Tensor,  # [C, N]
    image_width: int,
    image_height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C, tile_height, tile_width]
    flatten_ids: Tensor,  # [n_isects]
) -> Tuple[Tensor, Tensor, Tensor]:
    """Which 2D Gaussians of a depth slice hit which pixel -> (gaussian_ids [M], pixel_ids [M], camera_ids [M])."""
    C, N, _ = means2d.shape
    assert ray_transforms.shape == (C, N, 3, 3), ray_transforms.shape
    assert opacities.shape == (C, N), opacities.shape
    assert isect_offsets.shape[0] == C, isect_offsets.shape
    tile_height, tile_width = isect_offsets.shape[1:3]
    assert tile_height * tile_size >= image_height, f"Assert Failed: {tile_height} * {tile_size} >= {image_height}"
    assert tile_width * tile_size >= image_width, f"Assert Failed: {tile_width} * {tile_size} >= {image_width}"
    assert transmittances.shape == (C, image_height, image_width), transmittances.shape
    means2d, ray_transforms, opacities, transmittances = map(
        _f32, (means2d.detach(), ray_transforms.detach(), opacities.detach(), transmittances.detach()))
    isect_offsets, flatten_ids = isect_offsets.contiguous(), flatten_ids.contiguous()
    dev = means2d.device
    rs = int(min(max(range_start, 0), 0xFFFFFFFF))
    re = int(min(max(range_end, 0), 0xFFFFFFFF))
    geom = (rs, re, C, N, flatten_ids.numel(), ptr(means2d), ptr(ray_transforms), ptr(opacities), image_width, image_height,
            tile_size, tile_width, tile_height, ptr(isect_offsets), ptr(flatten_ids), ptr(transmittances))
    chunk_cnts = torch.empty((C, image_height, image_width), dtype=torch.int32, device=dev)
    with device_guard(dev):
        call("gsb200_indices_in_range_2dgs_count", *geom, ptr(chunk_cnts), stream_ptr(dev))
        ends = torch.cumsum(chunk_cnts.view(-1), 0, dtype=torch.int64)
        n_elems = int(ends[-1].item()) if ends.numel() else 0
        gaussian_ids = torch.empty(n_elems, dtype=torch.int64, device=dev)
        out_indices = torch.empty(n_elems, dtype=torch.int64, device=dev)
        if n_elems:
            chunk_starts = (ends - chunk_cnts.view(-1)).contiguous()
            call("gsb200_indices_in_range_2dgs_fill", *geom, ptr(chunk_starts), ptr(gaussian_ids), ptr(out_indices),
                 stream_ptr(dev))
    return gaussian_ids, out_indices % (image_width * image_height), out_indices // (image_width * image_height)


def accumulate_2dgs(
    means2d: Tensor,  # [C, N, 2]
    ray_transforms: Tensor,  # [C, N, 3, 3]
    opacities: Tensor,  # [C, N]
    colors: Tensor,  # [C, N, channels]
    normals: Tensor,  # [C, N, 3]
    gaussian_ids: Tensor,  # [M]
    pixel_ids: Tensor,  # [M]
    camera_ids: Tensor,  # [M]
    image_width: int,
    image_height: int,
) -> Tuple[Tensor, Tensor, Tensor]:
    """Differentiable alpha compositing of (pixel, Gaussian) pairs listed front to back per pixel -- the torch reference
    rasterizer of 2DGS (gsplat/cuda/_torch_impl_2dgs.py:78-168) without its nerfacc dependency: the exclusive cumulative
    transmittance per pixel is a segmented cumulative sum of log(1 - alpha).  -> (renders, alphas, normals)."""
    C, N = means2d.shape[:2]
    channels = colors.shape[-1]
    px = pixel_ids % image_width + 0.5
    py = pixel_ids // image_width + 0.5
    deltas = torch.stack([px, py], dim=-1) - means2d[camera_ids, gaussian_ids]
    M = ray_transforms[camera_ids, gaussian_ids]
    h_u = -M[..., 0, :3] + M[..., 2, :3] * px[..., None]
    h_v = -M[..., 1, :3] + M[..., 2, :3] * py[..., None]
    tmp = torch.cross(h_u, h_v, dim=-1)
    us, vs = tmp[..., 0] / tmp[..., 2], tmp[..., 1] / tmp[..., 2]
    sigmas = 0.5 * torch.minimum(us**2 + vs**2, 2 * (deltas[..., 0] ** 2 + deltas[..., 1] ** 2))
    alphas = torch.clamp_max(opacities[camera_ids, gaussian_ids] * torch.exp(-sigmas), 0.999)
    rays = camera_ids * image_height * image_width + pixel_ids
    total = C * image_height * image_width
    # exclusive per-ray product of (1 - alpha): the pairs of one ray are contiguous and ordered front to back
    # (float64: the running sum spans the whole list, the per-ray differences are tiny next to it)
    logt = torch.log1p(-alphas.double())
    csum = torch.cumsum(logt, 0)
    first = torch.ones_like(rays, dtype=torch.bool)
    first[1:] = rays[1:] != rays[:-1]
    start_idx = torch.nonzero(first)[:, 0]
    seg = torch.cumsum(first.to(torch.int64), 0) - 1
    before = (csum - logt)[start_idx][seg] if rays.numel() else csum
    weights = alphas * torch.exp(csum - logt - before).to(alphas.dtype)

    def acc(values, width):
        out = torch.zeros((total, width), dtype=weights.dtype, device=weights.device)
        return out.index_add_(0, rays, weights[:, None] * values)

    renders = acc(colors[camera_ids, gaussian_ids], channels).reshape(C, image_height, image_width, channels)
    out_alphas = acc(torch.ones_like(weights)[:, None], 1).reshape(C, image_height, image_width, 1)
    out_normals = acc(normals[camera_ids, gaussian_ids], 3).reshape(C, image_height, image_width, 3)
    return renders, out_alphas, out_normals
