"""`rasterization()` -- the render API of the hot path, same signature / returns / `meta` schema as the
reference's ``gsplat.rendering.rasterization`` (gsplat/rendering.py:28-582, meta keys :337-349 and
:512-525) so ``examples/simple_trainer*.py`` and ``profiling/main.py`` call it unchanged.

Orchestration differences from the reference (behaviour-preserving):
  * tile offsets come out of the per-tile intersection stage; there is no separate
    isect_offset_encode pass over the sorted keys;
  * for SH colours the camera position is read off the rigid view matrix (-R^T t) instead of a
    batched LU inverse of [C,4,4] (rendering.py:370); `torch.inverse` is used only if the matrix is
    not rigid;
  * distributed mode (view-parallel, attribute all-gather) lives in ``distributed.py``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
from typing_extensions import Literal

from .compression_simulation.ops import QuantOnLoad, solid
from .cuda._wrapper import (
    _isect_tiles_begin,
    _isect_tiles_with_offsets,
    fully_fused_projection,
    fully_fused_projection_quant,
    rasterize_to_pixels,
    spherical_harmonics,
    SH_SPLIT_MAX_REST,
    view_colors_from_sh,
    view_colors_from_sh_split,
)


def _camera_positions(viewmats: Tensor) -> Tensor:
    """World-space camera centres [C,3] of world-to-camera matrices [C,4,4]."""
    if viewmats.is_cuda and viewmats.dtype == torch.float32 and not (viewmats.requires_grad and torch.is_grad_enabled()):
        from ._cabi import call, device_guard, ptr, stream_ptr

        vm = viewmats.contiguous()
        campos = torch.empty((vm.shape[0], 3), dtype=torch.float32, device=vm.device)
        with device_guard(vm.device):
            call("gsb200_camera_positions", vm.shape[0], ptr(vm), ptr(campos), stream_ptr(vm.device))
        return campos
    R = viewmats[:, :3, :3]
    t = viewmats[:, :3, 3]
    # rigid transform: inverse(viewmat)[:3,3] = -R^T t ; differentiable w.r.t. viewmats
    return -(R.transpose(1, 2) @ t[..., None])[..., 0]


_colour_streams: Dict = {}


def _colour_stream(device) -> Optional["torch.cuda.Stream"]:
    """Side stream of the single-GPU colour stage (one per device), or None when disabled (GSB200_COLOR_STREAM=0)
    or not on a CUDA device.  The colour stage needs the projection's radii and nothing of the tile intersection, and
    the blend needs both: on its own stream the HBM-bound colour kernel runs while the intersection kernels sit in
    their atomics -- and, because autograd runs a node's backward on its forward's stream, the colour backward runs
    beside the projection backward."""
    import os

    if device.type != "cuda" or os.environ.get("GSB200_COLOR_STREAM", "1") == "0":
        return None
    st = _colour_streams.get(device.index)
    if st is None:
        st = _colour_streams[device.index] = torch.cuda.Stream(device=device)
        # the SH parameters' AccumulateGrad nodes live on the main stream and receive their gradient from the side
        # stream: intended (the engine synchronises the two), so silence torch's once-per-process notice about it
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
    return st


_inert_rows: Dict = {}


def _inert_row(widths: Tuple[int, ...], has_quats: bool, device) -> Tensor:
    """Attribute row of an inert padding splat: NaN mean (culled by the projection), identity quaternion, zero
    scale and opacity.  Cached per device: building it costs a synchronising pageable H2D copy."""
    key = (widths, has_quats, str(device))
    row = _inert_rows.get(key)
    if row is None:
        vals = [float("nan")] * 3 + [0.0] * (sum(widths) - 3)
        if has_quats:
            vals[3] = 1.0
        _inert_rows[key] = row = torch.tensor(vals, dtype=torch.float32, device=device)
    return row


def rasterization(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]
    scales: Tensor,  # [N, 3]
    opacities: Tensor,  # [N]
    colors,  # Tensor [(C,) N, D] or [(C,) N, K, 3]; or the pair (sh0 [N,1,3], shN [N,K-1,3]) -- additive
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: Literal["RGB", "D", "ED", "RGB+D", "RGB+ED"] = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: Literal["classic", "antialiased"] = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: Literal["pinhole", "ortho", "fisheye"] = "pinhole",
    covars: Optional[Tensor] = None,
    rigid_viewmats: bool = True,
) -> Tuple[Tensor, Tensor, Dict]:
    """Rasterize a set of 3D Gaussians (N) to a batch of image planes (C).

    Returns (render_colors [C,H,W,X], render_alphas [C,H,W,1], meta).  See the reference docstring
    (gsplat/rendering.py:55-222) for the meaning of every argument; `rigid_viewmats` is additive.

    Additive: with `sh_degree` set, `colors` may also be the pair ``(sh0, shN)`` of the trainers' two SH
    parameters ([N,1,3] and [N,K-1,3]) instead of ``torch.cat([sh0, shN], 1)`` (examples/simple_trainer.py:776).
    The dense single-process path then reads both arrays in place and writes their gradients directly (no
    [N,K,3] copy forward, no strided gradient copies backward); every other mode concatenates them first, so the
    result is the same as for the concatenated tensor.

    Additive (quantise on load): `quats`, `scales`, `opacities` and `sh0` of the pair may be
    ``compression_simulation.ops.QuantOnLoad`` objects -- what ``simulate_compression(..., fuse_into_render=True)``
    returns, possibly passed through ``torch.exp`` / ``torch.sigmoid``.  The dense single-process path then reads the RAW
    parameters and applies fake quantisation + activation inside the projection / colour kernels; every other mode
    materialises them first (same values).
    """
    meta: Dict = {}
    N = means.shape[0]
    C = viewmats.shape[0]
    device = means.device
    # quantise on load is served by the dense single-process kernels with fixed cameras; anything else sees tensors
    quant_on_load = (not distributed and not packed and covars is None
                     and not ((viewmats.requires_grad or Ks.requires_grad) and torch.is_grad_enabled()))
    if not quant_on_load:
        quats, scales, opacities = solid(quats), solid(scales), solid(opacities)
    quant_on_load = quant_on_load and any(isinstance(v, QuantOnLoad) for v in (quats, scales, opacities))
    means = solid(means)
    assert means.shape == (N, 3), means.shape
    if covars is None:
        assert quats.shape == (N, 4), quats.shape
        assert scales.shape == (N, 3), scales.shape
    else:
        assert covars.shape == (N, 3, 3), covars.shape
        quats, scales = None, None
        tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
        covars = covars[..., tri[0], tri[1]]
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    assert render_mode in ["RGB", "D", "ED", "RGB+D", "RGB+ED"], render_mode

    sh_pair, sh_pair_dist = None, None
    if isinstance(colors, (tuple, list)):
        assert sh_degree is not None and len(colors) == 2, "colors=(sh0, shN) needs sh_degree"
        sh0, shN = colors
        shN = solid(shN)
        if isinstance(sh0, QuantOnLoad) and (distributed or packed or shN.shape[1] > SH_SPLIT_MAX_REST or sh0.act is not None):
            sh0 = sh0.materialize()
        assert sh0.shape == (N, 1, 3), sh0.shape
        assert shN.dim() == 3 and shN.shape[0] == N and shN.shape[2] == 3, shN.shape
        assert (sh_degree + 1) ** 2 <= 1 + shN.shape[1], shN.shape
        if not distributed and not packed and shN.shape[1] <= SH_SPLIT_MAX_REST:
            sh_pair, colors = (sh0, shN), None
        elif distributed and not packed:
            sh_pair_dist, colors = (sh0, shN), None  # concatenated below only if the peer-memory path is unavailable
        else:
            colors = torch.cat([sh0, shN], dim=1)

    colors = solid(colors)
    if sh_pair is not None or sh_pair_dist is not None:
        pass
    elif sh_degree is None:
        assert (colors.dim() == 2 and colors.shape[0] == N) or (
            colors.dim() == 3 and colors.shape[:2] == (C, N)
        ), colors.shape
        if distributed:
            assert colors.dim() == 2, "Distributed mode only supports per-Gaussian colors."
    else:
        assert (colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3) or (
            colors.dim() == 4 and colors.shape[:2] == (C, N) and colors.shape[3] == 3
        ), colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[-2], colors.shape
        if distributed:
            assert colors.dim() == 3, "Distributed mode only supports per-Gaussian colors."

    if absgrad:
        assert not distributed, "AbsGrad is not supported in distributed mode."

    if distributed:
        # view-parallel: every rank owns a slice of the Gaussians and its own cameras; attributes are
        # all-gathered (autograd-aware: the backward is a reduce-scatter of attribute gradients) and
        # each rank then renders its cameras from the full set.  Replaces the reference's
        # project-everywhere + all-to-all exchange (rendering.py:282-294, 397-478); same API, gradients
        # reach the owner rank.
        from .distributed import gather_counts, start_gather

        counts = gather_counts(N, device)
        dist_world = len(counts)
        # two collectives each way: the geometry attributes travel as ONE [n, 11 | 10] tensor (every collective has
        # a fixed cost), the (large) colour / SH tensor separately -- it is only waited for after projection and
        # tile intersection have been enqueued, so its transfer overlaps with those kernels
        # Ragged shards are padded to a common length with INERT splats (NaN mean: culled by the projection, so
        # they never reach a tile, get zero gradients, and cost no compaction pass); the gathered set then has
        # world * max(N_r) rows, i.e. up to world - 1 more than sum(N_r).
        geo_parts = [means] + ([quats, scales] if covars is None else [covars]) + [opacities[:, None]]
        widths = [t.shape[1] for t in geo_parts]
        pad_row = _inert_row(tuple(widths), covars is None, device)
        # SH coefficients (the bulk of the bytes) are not gathered at all when the peer-memory path is usable: the
        # colour kernels pull the visible rows from their owners over NVLink, and the geometry all-gather is one
        # pull kernel on the same symmetric buffer (_peer.py)
        from . import _peer

        n_block = (max(counts) + 3) // 4 * 4
        means_local = means
        # Preferred path ("owner render"): nothing per-Gaussian moves at all.  Every rank projects ITS rows and evaluates
        # their SH colours for all views of the job and stores the results of the visible pairs into the viewers' tensors;
        # the backward sends the blend's gradients of those tensors home (_peer.owner_render_inputs).  Needs the trainers'
        # parameterisation: (quats, scales), the (sh0, shN) pair, classic rasterisation, fixed cameras.
        owner_render = (sh_pair_dist is not None and sh_degree is not None and not packed and covars is None
                        and rasterize_mode == "classic" and not sparse_grad
                        and sh_pair_dist[1].shape[1] <= SH_SPLIT_MAX_REST
                        and not ((viewmats.requires_grad or Ks.requires_grad) and torch.is_grad_enabled())
                        and _peer.owner_render_available(device, C, n_block))
        # Preferred colour path: the coefficients never move.  Every rank evaluates ITS rows for all views of the job and
        # stores the colours into the viewers' tensors; the backward sends colour gradients home (_peer.view_colors_owner)
        owner_colors = (not owner_render and sh_pair_dist is not None and sh_degree is not None and not packed
                        and sh_pair_dist[1].shape[1] <= SH_SPLIT_MAX_REST
                        and not (viewmats.requires_grad and torch.is_grad_enabled())
                        and _peer.owner_exchange_available(device, C, n_block))
        if sh_pair_dist is not None and not owner_colors:
            k_all = 1 + sh_pair_dist[1].shape[1]
            if _peer.available(device, k_all * 3, n_block):
                colors = sh_pair_dist  # published side by side, gradients come back per parameter
            else:
                colors = torch.cat(sh_pair_dist, dim=1)
        pull_colors = (not owner_colors and sh_degree is not None and not packed
                       and (isinstance(colors, tuple) or colors.dim() == 3)
                       and _peer.available(device, (1 + colors[1].shape[1] if isinstance(colors, tuple) else colors.shape[1]) * 3,
                                           n_block))
        colors_local = colors
        if owner_render:
            geo, pending_colors = None, None
        elif owner_colors:
            geo, published = _peer.publish_and_gather_geometry(geo_parts, pad_row, None, n_block)
            pending_colors = None
        elif pull_colors:
            geo, published = _peer.publish_and_gather_geometry(geo_parts, pad_row, colors_local, n_block)
            pending_colors = None
        else:
            pending_geo = start_gather(torch.cat(geo_parts, dim=1), counts, keep_padding=True, pad_row=pad_row)
            pending_colors = start_gather(colors, counts, keep_padding=True)
            geo = torch.split(pending_geo.finish(), widths, dim=1)
        if geo is not None:
            means = geo[0]
            if covars is None:
                quats, scales = geo[1], geo[2]
            else:
                covars = geo[1]
            opacities = geo[-1][:, 0]
            N = means.shape[0]
        else:
            N = dist_world * n_block
        colors = None
    else:
        pending_colors, pull_colors, owner_colors, owner_render = None, False, False, False

    if owner_render:
        from .cuda._wrapper import _CAMERA_MODELS

        radii, means2d, depths, conics, opacities, owner_colours = _peer.owner_render_inputs(
            sh_degree, means_local, quats, scales, opacities, sh_pair_dist, viewmats, Ks, width, height, n_block, eps2d,
            near_plane, far_plane, radius_clip, _CAMERA_MODELS[camera_model])
        proj_results = (radii, means2d, depths, conics, None)
    elif quant_on_load:
        *proj_results, opacities = fully_fused_projection_quant(
            means, quats, scales, opacities, viewmats, Ks, width, height, eps2d=eps2d, near_plane=near_plane,
            far_plane=far_plane, radius_clip=radius_clip, calc_compensations=(rasterize_mode == "antialiased"),
            camera_model=camera_model)
    else:
        proj_results = fully_fused_projection(
            means, covars, quats, scales, viewmats, Ks, width, height,
            eps2d=eps2d, packed=packed, near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip,
            sparse_grad=sparse_grad, calc_compensations=(rasterize_mode == "antialiased"), camera_model=camera_model,
        )
    if owner_render:
        camera_ids, gaussian_ids, compensations = None, None, None  # (opacities is already [C, N])
    elif packed:
        camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations = proj_results
        opacities = opacities[gaussian_ids]  # [nnz]
    else:
        radii, means2d, depths, conics, compensations = proj_results
        opacities = opacities.unsqueeze(0) if C == 1 else opacities.repeat(C, 1)  # [C, N] (a view for one camera)
        camera_ids, gaussian_ids = None, None
    if compensations is not None:
        opacities = opacities * compensations

    meta.update({
        "camera_ids": camera_ids, "gaussian_ids": gaussian_ids, "radii": radii, "means2d": means2d,
        "depths": depths, "conics": conics, "opacities": opacities,
    })

    # Tile intersection needs only the projection outputs: enqueue it before the colour stage so that, in
    # distributed mode, the colour all-gather keeps overlapping with it.
    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))
    # single-GPU SH colours from per-Gaussian coefficients: on a side stream that forks here, after the projection
    # and BEFORE the intersection kernels are enqueued, and joins the main stream in front of the blend
    side, fork = None, None
    if (sh_degree is not None and not owner_render and not packed and pending_colors is None and not owner_colors
            and not pull_colors and (sh_pair is not None or colors.dim() == 3)):
        side = _colour_stream(means.device)
        if side is not None:
            campos_side = _camera_positions(viewmats) if rigid_viewmats else torch.inverse(viewmats)[:, :3, 3]
            fork = torch.cuda.Event()
            fork.record(torch.cuda.current_stream(means.device))
    isect_pending = _isect_tiles_begin(
        means2d, radii, depths, tile_size, tile_width, tile_height,
        packed=packed, n_cameras=C, camera_ids=camera_ids, gaussian_ids=gaussian_ids,
    )
    if pending_colors is not None:
        colors = pending_colors.finish()

    if owner_render:
        colors = owner_colours
    elif sh_degree is None:
        if packed:
            colors = colors[gaussian_ids] if colors.dim() == 2 else colors[camera_ids, gaussian_ids]
        else:
            colors = colors.expand(C, -1, -1) if colors.dim() == 2 else colors
    elif side is not None:
        main = torch.cuda.current_stream(means.device)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            if sh_pair is not None:
                colors = view_colors_from_sh_split(sh_degree, means, campos_side, sh_pair[0], sh_pair[1], radii)
            else:
                colors = view_colors_from_sh(sh_degree, means, campos_side, colors, radii)
        colors.record_stream(main)  # allocated from the side stream's pool, read by the blend on the main stream
    else:
        campos = _camera_positions(viewmats) if rigid_viewmats else torch.inverse(viewmats)[:, :3, 3]
        if sh_pair is not None:
            colors = view_colors_from_sh_split(sh_degree, means, campos, sh_pair[0], sh_pair[1], radii)
        elif owner_colors:
            colors = _peer.view_colors_owner(sh_degree, means_local, campos, sh_pair_dist, radii, n_block)
        elif pull_colors:
            colors = _peer.view_colors_from_sh_sharded(sh_degree, means, campos, colors_local, radii, n_block, published)
        elif not packed and colors.dim() == 3:
            # dense layout, per-Gaussian coefficients: one fused kernel (dirs, SH, +0.5, clamp)
            colors = view_colors_from_sh(sh_degree, means, campos, colors, radii)
        elif packed:
            dirs = means[gaussian_ids, :] - campos[camera_ids]  # [nnz, 3]
            masks = radii > 0
            shs = colors[gaussian_ids, :, :] if colors.dim() == 3 else colors[camera_ids, gaussian_ids, :, :]
            colors = torch.clamp_min(spherical_harmonics(sh_degree, dirs, shs, masks=masks) + 0.5, 0.0)
        else:
            dirs = means[None, :, :] - campos[:, None, :]  # [C, N, 3]
            masks = radii > 0
            colors = torch.clamp_min(spherical_harmonics(sh_degree, dirs, colors, masks=masks) + 0.5, 0.0)

    # the host waits for the intersection count only now: the colour kernels above are already enqueued behind the
    # scatter + sort, so the device has work for the whole wait and for the launch of the blend that follows
    tiles_per_gauss, isect_ids, flatten_ids, isect_offsets = isect_pending.finish()
    if side is not None:
        torch.cuda.current_stream(means.device).wait_stream(side)

    if render_mode in ["RGB+D", "RGB+ED"]:
        colors = torch.cat((colors, depths[..., None]), dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1, device=backgrounds.device)], dim=-1)
    elif render_mode in ["D", "ED"]:
        colors = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1, device=backgrounds.device)

    meta.update({
        "tile_width": tile_width, "tile_height": tile_height, "tiles_per_gauss": tiles_per_gauss,
        "isect_ids": isect_ids, "flatten_ids": flatten_ids, "isect_offsets": isect_offsets,
        "width": width, "height": height, "tile_size": tile_size, "n_cameras": C,
    })

    if colors.shape[-1] > channel_chunk:
        n_chunks = (colors.shape[-1] + channel_chunk - 1) // channel_chunk
        render_colors, render_alphas = [], []
        for i in range(n_chunks):
            colors_chunk = colors[..., i * channel_chunk:(i + 1) * channel_chunk]
            backgrounds_chunk = (
                backgrounds[..., i * channel_chunk:(i + 1) * channel_chunk] if backgrounds is not None else None
            )
            rc, ra = rasterize_to_pixels(
                means2d, conics, colors_chunk, opacities, width, height, tile_size, isect_offsets, flatten_ids,
                backgrounds=backgrounds_chunk, packed=packed, absgrad=absgrad,
            )
            render_colors.append(rc)
            render_alphas.append(ra)
        render_colors = torch.cat(render_colors, dim=-1)
        render_alphas = render_alphas[0]
    else:
        render_colors, render_alphas = rasterize_to_pixels(
            means2d, conics, colors, opacities, width, height, tile_size, isect_offsets, flatten_ids,
            backgrounds=backgrounds, packed=packed, absgrad=absgrad,
        )
    if render_mode in ["ED", "RGB+ED"]:
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1
        )
    return render_colors, render_alphas, meta


# (the reference defines these in rendering.py as well)
from .rendering_2dgs import rasterization_2dgs, rasterization_2dgs_inria_wrapper  # noqa: E402,F401
