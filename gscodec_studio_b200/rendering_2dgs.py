"""`rasterization_2dgs()` -- the render API of the 2D Gaussian (surfel) family, same signature / returns / `meta` keys as
the reference's ``gsplat.rendering.rasterization_2dgs`` (gsplat/rendering.py:1002-1319), on the surfel kernels of
csrc/surfel.cu and the tile-intersection / SH stages of the 3DGS path.

`rasterization_2dgs_inria_wrapper` (rendering.py:1322-1479) drives a third-party rasterizer
(hbb1/diff-surfel-rasterization), which is not part of the reference repository: not provided.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
from typing_extensions import Literal

from .cuda._wrapper import _isect_tiles_with_offsets, spherical_harmonics
from .cuda._wrapper_2dgs import fully_fused_projection_2dgs, rasterize_to_pixels_2dgs
from .utils import depth_to_normal


def rasterization_2dgs(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]
    scales: Tensor,  # [N, 3]
    opacities: Tensor,  # [N]
    colors: Tensor,  # [(C,) N, D] or [(C,) N, K, 3]
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = False,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: Literal["RGB", "D", "ED", "RGB+D", "RGB+ED"] = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    distloss: bool = False,
    depth_mode: Literal["expected", "median"] = "expected",
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Dict]:
    """Rasterize a set of 2D Gaussians (N) to a batch of image planes (C).

    Returns (render_colors [C,H,W,X], render_alphas [C,H,W,1], render_normals [C,H,W,3] (world space),
    normals_from_depth [C,H,W,3] or None, render_distort [C,H,W,1], render_median [C,H,W,1], meta)."""
    N = means.shape[0]
    C = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    assert render_mode in ["RGB", "D", "ED", "RGB+D", "RGB+ED"], render_mode
    if sh_degree is None:
        assert (colors.dim() == 2 and colors.shape[0] == N) or (colors.dim() == 3 and colors.shape[:2] == (C, N)), colors.shape
    else:
        assert colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3, colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[1], colors.shape
    if distloss:
        assert render_mode in ["D", "ED", "RGB+D", "RGB+ED"], "distloss requires depth rendering"

    proj = fully_fused_projection_2dgs(means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                                       radius_clip, packed, sparse_grad)
    if packed:
        camera_ids, gaussian_ids, radii, means2d, depths, ray_transforms, normals = proj
        opacities = opacities[gaussian_ids]
    else:
        radii, means2d, depths, ray_transforms, normals = proj
        opacities = opacities.repeat(C, 1)
        camera_ids, gaussian_ids = None, None
    # dummy input of the blend whose gradient carries the densification signal of 2DGS (rendering.py:1170-1172)
    densify = torch.zeros_like(means2d, dtype=means.dtype, requires_grad=True)

    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))
    tiles_per_gauss, isect_ids, flatten_ids, isect_offsets = _isect_tiles_with_offsets(
        means2d, radii, depths, tile_size, tile_width, tile_height, packed=packed, n_cameras=C, camera_ids=camera_ids,
        gaussian_ids=gaussian_ids)

    if not (colors.dim() == 3 and sh_degree is None):
        colors = colors[gaussian_ids] if packed else colors.expand(C, *([-1] * colors.dim()))
    elif packed:
        colors = colors[camera_ids, gaussian_ids, :]
    if sh_degree is not None:
        campos = torch.inverse(viewmats)[:, :3, 3]
        dirs = means[gaussian_ids, :] - campos[camera_ids] if packed else means[None, :, :] - campos[:, None, :]
        colors = torch.clamp_min(spherical_harmonics(sh_degree, dirs, colors, masks=radii > 0) + 0.5, 0.0)
    if render_mode in ["RGB+D", "RGB+ED"]:
        colors = torch.cat((colors, depths[..., None]), dim=-1)
    elif render_mode in ["D", "ED"]:
        colors = depths[..., None]

    render_colors, render_alphas, render_normals, render_distort, render_median = rasterize_to_pixels_2dgs(
        means2d, ray_transforms, colors, opacities, normals, densify, width, height, tile_size, isect_offsets, flatten_ids,
        backgrounds=backgrounds, packed=packed, absgrad=absgrad, distloss=distloss)

    normals_from_depth = None
    if render_mode in ["ED", "RGB+ED"]:
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1)
    if render_mode in ["RGB+ED", "RGB+D"]:
        depth_for_normal = render_colors[..., -1:] if depth_mode == "expected" else render_median
        normals_from_depth = depth_to_normal(depth_for_normal, torch.linalg.inv(viewmats), Ks).squeeze(0)

    meta = {
        "camera_ids": camera_ids, "gaussian_ids": gaussian_ids, "radii": radii, "means2d": means2d, "depths": depths,
        "ray_transforms": ray_transforms, "opacities": opacities, "normals": normals,
        "tile_width": tile_width, "tile_height": tile_height, "tiles_per_gauss": tiles_per_gauss, "isect_ids": isect_ids,
        "flatten_ids": flatten_ids, "isect_offsets": isect_offsets, "width": width, "height": height, "tile_size": tile_size,
        "n_cameras": C, "render_distort": render_distort,
        "gradient_2dgs": densify,  # its .grad is what the 2DGS densification strategy reads
    }
    # camera space -> world space
    render_normals = torch.einsum("...ij,...hwj->...hwi", torch.linalg.inv(viewmats)[..., :3, :3], render_normals)
    return render_colors, render_alphas, render_normals, normals_from_depth, render_distort, render_median, meta


def rasterization_2dgs_inria_wrapper(*args, **kwargs):
    raise NotImplementedError(
        "rasterization_2dgs_inria_wrapper drives the third-party diff_surfel_rasterization backend "
        "(github.com/hbb1/diff-surfel-rasterization), which is neither part of the reference repository nor of this package; "
        "use rasterization_2dgs")
