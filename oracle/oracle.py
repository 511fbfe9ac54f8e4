"""CPU oracle for the gsplat rasterization hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``gscodec_studio_b200/`` imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do,
and there only as the checker or the reported CPU baseline, never as the product.

This is a *restatement* (written from the reference's behaviour, not copied) of

* stage 1  projection          gsplat/cuda/csrc/fully_fused_projection_fwd.cu:45-196 (dense),
                               fully_fused_projection_packed_fwd.cu:183-196 (packed radius rule),
                               cross-checked with gsplat/cuda/_torch_impl.py:250-327
* colour   spherical harmonics gsplat/cuda/include/spherical_harmonics.cuh:13-136,
                               gsplat/rendering.py:352-392
* stage 2  tile intersection   gsplat/cuda/csrc/isect_tiles.cu:16-104,308-354,
                               cross-checked with _torch_impl.py:330-429
* stage 3/4a blend fwd/bwd     oracle_cpu.c (rasterize_to_pixels_{fwd,bwd}.cu)
* stage 4b projection / SH backward: torch autograd through the float32 restatement
* quant-sim                    gsplat/compression_simulation/ops.py:39-75,
                               entropy_model.py:195-254,347-390

Pinning (see tests/test_oracle_golden.py): projection / SH / isect / offsets / quant are pinned
against the reference's own ``_torch_impl.py`` and ``ops.py`` imported by file path in the build
container (vectors under tests/golden/, generator oracle/make_golden.py); the blend is pinned
against the unmodified reference CUDA kernels run on a B200 (oracle/make_golden_gpu.py).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle_cpu.c -> liboracle.so (gcc only)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle_cpu.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"]
        )
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads() -> int:
    return int(_lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    _lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def _p(t: Optional[Tensor]):
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


# --------------------------------------------------------------------------------------
# stage 1: projection
# --------------------------------------------------------------------------------------
def quat_to_rotmat(quats: Tensor) -> Tensor:
    """wxyz quaternion (normalised inside, quat.cuh:9-33) -> rotation matrices [...,3,3]."""
    q = quats * torch.rsqrt((quats * quats).sum(-1, keepdim=True))
    w, x, y, z = q.unbind(-1)
    rows = [
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
    ]
    return torch.stack(rows, -1).reshape(quats.shape[:-1] + (3, 3))


def quat_scale_to_covar(quats: Tensor, scales: Tensor) -> Tensor:
    """Sigma = (R diag(s)) (R diag(s))^T  (quat_scale_to_covar_preci.cuh:17-24)."""
    M = quat_to_rotmat(quats) * scales[..., None, :]
    return M @ M.transpose(-1, -2)


def covars_from_triu(c6: Tensor) -> Tensor:
    xx, xy, xz, yy, yz, zz = c6.unbind(-1)
    return torch.stack([xx, xy, xz, xy, yy, yz, xz, yz, zz], -1).reshape(c6.shape[:-1] + (3, 3))


def projection(
    means: Tensor,  # [N,3]
    covars: Optional[Tensor],  # [N,6] triu or None
    quats: Optional[Tensor],  # [N,4]
    scales: Optional[Tensor],  # [N,3]
    viewmats: Tensor,  # [C,4,4]
    Ks: Tensor,  # [C,3,3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    calc_compensations: bool = False,
    camera_model: str = "pinhole",
    packed_radius_rule: bool = False,
):
    """Dense projection, differentiable (autograd = stage 4b oracle).

    Returns radii i32 [C,N], means2d [C,N,2], depths [C,N], conics [C,N,3], compensations|None.
    Entries with radii == 0 are culled; their other outputs are zero-filled here (the reference
    leaves them uninitialised, fully_fused_projection_fwd.cu:236-240).
    """
    C, N = viewmats.shape[0], means.shape[0]
    R = viewmats[:, :3, :3]
    t = viewmats[:, :3, 3]
    mean_c = torch.einsum("cij,nj->cni", R, means) + t[:, None, :]
    Sigma = covars_from_triu(covars) if covars is not None else quat_scale_to_covar(quats, scales)
    Sigma_c = torch.einsum("cij,njk,clk->cnil", R, Sigma, R)
    x, y, z = mean_c.unbind(-1)
    fx, fy = Ks[:, 0, 0, None], Ks[:, 1, 1, None]
    cx, cy = Ks[:, 0, 2, None], Ks[:, 1, 2, None]
    O = torch.zeros_like(z)
    if camera_model == "pinhole":  # proj.cuh:80-119
        tan_fovx, tan_fovy = 0.5 * width / fx, 0.5 * height / fy
        lim_xp, lim_xn = (width - cx) / fx + 0.3 * tan_fovx, cx / fx + 0.3 * tan_fovx
        lim_yp, lim_yn = (height - cy) / fy + 0.3 * tan_fovy, cy / fy + 0.3 * tan_fovy
        rz = 1.0 / z
        tx = z * torch.minimum(lim_xp.expand_as(z), torch.maximum(-lim_xn.expand_as(z), x * rz))
        ty = z * torch.minimum(lim_yp.expand_as(z), torch.maximum(-lim_yn.expand_as(z), y * rz))
        J = torch.stack([fx * rz, O, -fx * tx * rz * rz, O, fy * rz, -fy * ty * rz * rz], -1)
        means2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], -1)
    elif camera_model == "ortho":  # proj.cuh:9-36
        J = torch.stack([fx + O, O, O, O, fy + O, O], -1)
        means2d = torch.stack([fx * x + cx, fy * y + cy], -1)
    elif camera_model == "fisheye":  # proj.cuh:202-241
        eps = 1e-7
        xy_len = torch.sqrt(x * x + y * y) + eps
        theta = torch.atan2(xy_len, z + eps)
        means2d = torch.stack([x * fx * theta / xy_len + cx, y * fy * theta / xy_len + cy], -1)
        x2, y2, xy = x * x + eps, y * y, x * y
        x2y2 = x2 + y2
        inv = 1.0 / (x2y2 + z * z)
        b = torch.atan2(xy_len, z) / xy_len / x2y2
        a = z * inv / x2y2
        J = torch.stack(
            [fx * (x2 * a + y2 * b), fx * xy * (a - b), -fx * x * inv,
             fy * xy * (a - b), fy * (y2 * a + x2 * b), -fy * y * inv], -1)
    else:
        raise ValueError(camera_model)
    J = J.reshape(C, N, 2, 3)
    cov2d = J @ Sigma_c @ J.transpose(-1, -2)
    det_orig = cov2d[..., 0, 0] * cov2d[..., 1, 1] - cov2d[..., 0, 1] * cov2d[..., 1, 0]
    c00 = cov2d[..., 0, 0] + eps2d
    c11 = cov2d[..., 1, 1] + eps2d
    c01 = cov2d[..., 0, 1]
    det = c00 * c11 - c01 * cov2d[..., 1, 0]
    valid = (z >= near_plane) & (z <= far_plane) & (det > 0)
    det_safe = torch.where(det > 0, det, torch.ones_like(det))
    comp = torch.sqrt(torch.clamp(det_orig / det_safe, min=0.0))
    inv_det = 1.0 / det_safe
    conics = torch.stack([c11 * inv_det, -c01 * inv_det, c00 * inv_det], -1)
    with torch.no_grad():
        b = 0.5 * (c00 + c11)
        if packed_radius_rule:
            s = torch.sqrt(torch.clamp(b * b - det, min=0.1))
            radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(b + s, b - s)))
        else:
            radius = torch.ceil(3.0 * torch.sqrt(b + torch.sqrt(torch.clamp(b * b - det, min=0.01))))
        valid = valid & ~(radius <= radius_clip)
        mx, my = means2d[..., 0], means2d[..., 1]
        outside = (mx + radius <= 0) | (mx - radius >= width) | (my + radius <= 0) | (my - radius >= height)
        valid = valid & ~outside
        radii = torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32)
    vm = valid
    means2d = torch.where(vm[..., None], means2d, torch.zeros_like(means2d))
    depths = torch.where(vm, z, torch.zeros_like(z))
    conics = torch.where(vm[..., None], conics, torch.zeros_like(conics))
    comp = torch.where(vm, comp, torch.zeros_like(comp)) if calc_compensations else None
    return radii, means2d, depths, conics, comp


# --------------------------------------------------------------------------------------
# colour: spherical harmonics (real SH, Sloan's polynomial forms, degree <= 4)
# --------------------------------------------------------------------------------------
def sh_basis(degree: int, dirs: Tensor) -> Tensor:
    """[..., 3] (un-normalised) -> [..., (degree+1)^2] basis values (spherical_harmonics.cuh:21-94)."""
    d = dirs * torch.rsqrt((dirs * dirs).sum(-1, keepdim=True))
    x, y, z = d.unbind(-1)
    Y = [torch.full_like(x, 0.2820947917738781)]
    if degree >= 1:
        Y += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
    if degree >= 2:
        z2 = z * z
        c1, s1 = x * x - y * y, 2.0 * x * y
        t0 = -1.092548430592079 * z
        Y6 = 0.9461746957575601 * z2 - 0.3153915652525201
        Y += [0.5462742152960395 * s1, t0 * y, Y6, t0 * x, 0.5462742152960395 * c1]
    if degree >= 3:
        t1 = -2.285228997322329 * z2 + 0.4570457994644658
        t2 = 1.445305721320277 * z
        c2, s2 = x * c1 - y * s1, x * s1 + y * c1
        Y12 = z * (1.865881662950577 * z2 - 1.119528997770346)
        Y += [-0.5900435899266435 * s2, t2 * s1, t1 * y, Y12, t1 * x, t2 * c1, -0.5900435899266435 * c2]
    if degree >= 4:
        u0 = z * (-4.683325804901025 * z2 + 2.007139630671868)
        u1 = 3.31161143515146 * z2 - 0.47308734787878
        u2 = -1.770130769779931 * z
        c3, s3 = x * c2 - y * s2, x * s2 + y * c2
        Y20 = 1.984313483298443 * z * Y12 - 1.006230589874905 * Y6
        Y += [0.6258357354491763 * s3, u2 * s2, u1 * s1, u0 * y, Y20, u0 * x, u1 * c1, u2 * c2,
              0.6258357354491763 * c3]
    return torch.stack(Y, -1)


def spherical_harmonics(degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """colors[..., 3] = sum_k Y_k(dir) * coeffs[..., k, :]; masked-out entries are zero."""
    K = (degree + 1) ** 2
    Y = sh_basis(degree, dirs)
    out = (Y[..., None] * coeffs[..., :K, :]).sum(-2)
    if masks is not None:
        out = torch.where(masks[..., None], out, torch.zeros_like(out))
    return out


# --------------------------------------------------------------------------------------
# stage 2: tile intersection, stable sort, offsets
# --------------------------------------------------------------------------------------
def _bit_width(v: int) -> int:
    return int(math.floor(math.log2(v))) + 1  # isect_tiles.cu:155-156


def isect_tiles(
    means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int, tile_height: int,
    n_cameras: int, camera_ids: Optional[Tensor] = None, sort: bool = True,
) -> Tuple[Tensor, Tensor, Tensor]:
    """-> tiles_per_gauss i32 (same shape as radii), isect_ids i64 [I], flatten_ids i32 [I]."""
    shape = radii.shape
    m = means2d.detach().reshape(-1, 2).numpy().astype(np.float32)
    r = radii.reshape(-1).numpy().astype(np.float32)
    d = depths.detach().reshape(-1).numpy().astype(np.float32)
    n = r.shape[0]
    ts = np.float32(tile_size)
    tr, tx, ty = r / ts, m[:, 0] / ts, m[:, 1] / ts

    def clampu(v, hi):  # float -> uint32 saturating cast, then min(max(0, .), hi)  (isect_tiles.cu:66-71)
        v = np.nan_to_num(v, nan=0.0, posinf=4e9, neginf=0.0)
        return np.clip(v, 0, hi).astype(np.int64)

    x0, x1 = clampu(np.floor(tx - tr), tile_width), clampu(np.ceil(tx + tr), tile_width)
    y0, y1 = clampu(np.floor(ty - tr), tile_height), clampu(np.ceil(ty + tr), tile_height)
    nx, ny = x1 - x0, y1 - y0
    tpg = np.where(r > 0, nx * ny, 0).astype(np.int64)
    total = int(tpg.sum())
    tile_bits = _bit_width(tile_width * tile_height)
    cam_bits = _bit_width(n_cameras)
    if camera_ids is not None:
        cid = camera_ids.reshape(-1).numpy().astype(np.int64)
    else:
        N = n // n_cameras
        cid = np.arange(n, dtype=np.int64) // max(N, 1)
    # expand: emission order = ascending flatten index, then row-major (y, x)  (isect_tiles.cu:89-103)
    owner = np.repeat(np.arange(n, dtype=np.int64), tpg)
    start = np.cumsum(tpg) - tpg
    local = np.arange(total, dtype=np.int64) - start[owner]
    nxo = np.maximum(nx[owner], 1)
    tyy = y0[owner] + local // nxo
    txx = x0[owner] + local % nxo
    tile_id = tyy * tile_width + txx
    depth_bits = d.view(np.int32).astype(np.int64)  # sign-extending, like the reference's cast
    keys = (cid[owner] << (32 + tile_bits)) | (tile_id << 32) | depth_bits[owner]
    flat = owner.astype(np.int32)
    if sort and total:
        mask = (np.int64(1) << (32 + tile_bits + cam_bits)) - 1 if 32 + tile_bits + cam_bits < 63 else np.int64(-1)
        order = np.argsort(keys & mask, kind="stable")
        keys, flat = keys[order], flat[order]
    return (
        torch.from_numpy(tpg.astype(np.int32)).reshape(shape),
        torch.from_numpy(keys.astype(np.int64)),
        torch.from_numpy(flat),
    )


def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """Start index of every (camera, tile) run in the sorted key list (isect_tiles.cu:308-354)."""
    n_tiles = tile_width * tile_height
    tile_bits = _bit_width(n_tiles)
    ids = isect_ids.numpy() >> 32
    flat = (ids >> tile_bits) * n_tiles + (ids & ((1 << tile_bits) - 1))
    counts = np.bincount(flat, minlength=n_cameras * n_tiles)[: n_cameras * n_tiles]
    off = np.cumsum(counts) - counts
    return torch.from_numpy(off.astype(np.int32)).reshape(n_cameras, tile_height, tile_width)


# --------------------------------------------------------------------------------------
# stage 3 / 4a: blend (C)
# --------------------------------------------------------------------------------------
def blend_fwd(means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size,
              isect_offsets, flatten_ids, tile_range: Optional[Tuple[int, int]] = None):
    """CPU stage 3.  means2d [C,N,2]|[nnz,2] ... -> render_colors [C,H,W,D], alphas [C,H,W,1], last_ids."""
    C, th, tw = isect_offsets.shape
    D = colors.shape[-1]
    f = lambda t: None if t is None else t.detach().contiguous().float()
    means2d, conics, colors, opacities, backgrounds = map(f, (means2d, conics, colors, opacities, backgrounds))
    masks_u8 = None if masks is None else masks.contiguous().to(torch.uint8)
    offs = isect_offsets.contiguous().int()
    fids = flatten_ids.contiguous().int()
    out_c = torch.zeros(C, height, width, D)
    out_a = torch.zeros(C, height, width, 1)
    last = torch.zeros(C, height, width, dtype=torch.int32)
    t0, t1 = tile_range if tile_range is not None else (0, C * th * tw)
    _lib().oracle_blend_fwd(
        C, fids.numel(), D, _p(means2d), _p(conics), _p(colors), _p(opacities), _p(backgrounds),
        _p(masks_u8), width, height, tile_size, tw, th, _p(offs), _p(fids), _p(out_c), _p(out_a),
        _p(last), t0, t1)
    return out_c, out_a, last


def blend_bwd(means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size,
              isect_offsets, flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas,
              absgrad: bool = False, tile_range: Optional[Tuple[int, int]] = None):
    """CPU stage 4a -> (v_means2d_abs|None, v_means2d, v_conics, v_colors, v_opacities)."""
    C, th, tw = isect_offsets.shape
    D = colors.shape[-1]
    f = lambda t: None if t is None else t.detach().contiguous().float()
    means2d, conics, colors, opacities, backgrounds, render_alphas, v_render_colors, v_render_alphas = map(
        f, (means2d, conics, colors, opacities, backgrounds, render_alphas, v_render_colors, v_render_alphas))
    masks_u8 = None if masks is None else masks.contiguous().to(torch.uint8)
    offs = isect_offsets.contiguous().int()
    fids = flatten_ids.contiguous().int()
    last_ids = last_ids.contiguous().int()
    v_abs = torch.zeros_like(means2d) if absgrad else None
    v_m, v_c, v_col, v_o = (torch.zeros_like(means2d), torch.zeros_like(conics), torch.zeros_like(colors),
                            torch.zeros_like(opacities))
    t0, t1 = tile_range if tile_range is not None else (0, C * th * tw)
    _lib().oracle_blend_bwd(
        C, fids.numel(), D, _p(means2d), _p(conics), _p(colors), _p(opacities), _p(backgrounds),
        _p(masks_u8), width, height, tile_size, tw, th, _p(offs), _p(fids), _p(render_alphas),
        _p(last_ids), _p(v_render_colors), _p(v_render_alphas), _p(v_abs), _p(v_m), _p(v_c),
        _p(v_col), _p(v_o), t0, t1)
    return v_abs, v_m, v_c, v_col, v_o


class _Blend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size,
                isect_offsets, flatten_ids):
        rc, ra, last = blend_fwd(means2d, conics, colors, opacities, backgrounds, masks, width, height,
                                 tile_size, isect_offsets, flatten_ids)
        ctx.save_for_backward(means2d, conics, colors, opacities, backgrounds, masks, isect_offsets,
                              flatten_ids, ra, last)
        ctx.dims = (width, height, tile_size)
        return rc, ra

    @staticmethod
    def backward(ctx, v_rc, v_ra):
        means2d, conics, colors, opacities, backgrounds, masks, offs, fids, ra, last = ctx.saved_tensors
        width, height, tile_size = ctx.dims
        _, v_m, v_c, v_col, v_o = blend_bwd(means2d, conics, colors, opacities, backgrounds, masks, width,
                                            height, tile_size, offs, fids, ra, last, v_rc.contiguous(),
                                            v_ra.contiguous())
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[4]:
            v_bg = (v_rc * (1.0 - ra)).sum(dim=(1, 2))  # _wrapper.py:1008-1011
        return v_m, v_c, v_col, v_o, v_bg, None, None, None, None, None, None


def rasterize_to_pixels(means2d, conics, colors, opacities, width, height, tile_size, isect_offsets,
                        flatten_ids, backgrounds=None, masks=None):
    """Differentiable CPU blend (stage 3 + 4a)."""
    return _Blend.apply(means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size,
                        isect_offsets, flatten_ids)


# --------------------------------------------------------------------------------------
# whole path: rasterization() on CPU (dense layout; gsplat/rendering.py:225-582, 1 process)
# --------------------------------------------------------------------------------------
def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                  near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, sh_degree=None,
                  tile_size=16, backgrounds=None, render_mode="RGB", rasterize_mode="classic",
                  camera_model="pinhole", covars=None, packed=False):
    C, N = viewmats.shape[0], means.shape[0]
    if covars is not None:
        idx = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
        covars = covars[..., idx[0], idx[1]]
        quats = scales = None
    radii, means2d, depths, conics, comp = projection(
        means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
        radius_clip, rasterize_mode == "antialiased", camera_model, packed_radius_rule=packed)
    opac = opacities[None, :].expand(C, N)
    if comp is not None:
        opac = opac * comp
    if sh_degree is None:
        cols = colors.expand(C, -1, -1) if colors.dim() == 2 else colors
    else:
        campos = torch.inverse(viewmats)[:, :3, 3]
        dirs = means[None] - campos[:, None]
        shs = colors.expand(C, -1, -1, -1) if colors.dim() == 3 else colors
        cols = torch.clamp_min(spherical_harmonics(sh_degree, dirs, shs, radii > 0) + 0.5, 0.0)
    if render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, depths[..., None]], -1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1)], -1)
    elif render_mode in ("D", "ED"):
        cols = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1)
    tw, th = math.ceil(width / tile_size), math.ceil(height / tile_size)
    tpg, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tw, th, C)
    offsets = isect_offset_encode(isect_ids, C, tw, th)
    rc, ra = rasterize_to_pixels(means2d, conics, cols.contiguous(), opac.contiguous(), width, height,
                                 tile_size, offsets, flatten_ids, backgrounds)
    if render_mode in ("ED", "RGB+ED"):
        rc = torch.cat([rc[..., :-1], rc[..., -1:] / ra.clamp(min=1e-10)], -1)
    meta = dict(radii=radii, means2d=means2d, depths=depths, conics=conics, opacities=opac,
                tiles_per_gauss=tpg, isect_ids=isect_ids, flatten_ids=flatten_ids, isect_offsets=offsets,
                tile_width=tw, tile_height=th, width=width, height=height, tile_size=tile_size, n_cameras=C)
    return rc, ra, meta


# --------------------------------------------------------------------------------------
# quant-sim (compression_simulation/ops.py:39-75) and factorized entropy model
# --------------------------------------------------------------------------------------
def fake_quantize_noise(x: Tensor, lo: float, hi: float, bitwidth: int, noise: Tensor):
    """y = clamp(x, lo, hi) + noise * q_step, noise ~ U(-.5,.5) supplied by the caller (ops.py:44-47)."""
    q_step = (hi - lo) / (2 ** bitwidth - 1)
    return torch.clamp(x, lo, hi) + noise * q_step, q_step


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bitwidth, lo, hi):
        x = x.clamp_(lo, hi)  # in place on the parameter, ops.py:63
        n = (x - lo) / (hi - lo)
        qs = 1 / (2 ** bitwidth - 1)
        return (n / qs).round() * qs * (hi - lo) + lo

    @staticmethod
    def backward(ctx, g):
        return g, None, None, None


def fake_quantize_round(x: Tensor, lo: float, hi: float, bitwidth: int):
    q_step = (hi - lo) / (2 ** bitwidth - 1)
    return _RoundSTE.apply(x, bitwidth, lo, hi), q_step


class _LowerBound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x)
        ctx.bound = bound
        return torch.clamp_min(x, bound)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ((x >= ctx.bound) | (g < 0)) * g, None


def entropy_param_index(n: int, ch: int, N: int) -> np.ndarray:
    """Parameter set used for element (n, c=?, half=?) in the reference's `_refactor` forward.

    The reference pads N to Npad = N + (32 - N % 32), views [2ch,1,Npad] as [64ch,1,Npad/32] and
    tiles the weights with repeat(64,1,1) (entropy_model.py:218-243), so row r = half*ch + c and
    chunk j = n // (Npad/32) evaluate with parameter set (r*32 + j) % ch.  Returned: [2, ch, N] ints.
    """
    Npad = N + (32 - N % 32)
    L = Npad // 32
    r = np.arange(2 * ch)[:, None]
    j = (np.arange(N) // L)[None, :]
    return ((r * 32 + j) % ch).reshape(2, ch, N)


def entropy_factorized_bits(x: Tensor, q: float, matrices, biases, factors, likelihood_bound=1e-6) -> Tensor:
    """bits[N,ch] of the factorized prior (entropy_model.py:201-254), elementwise restatement.

    matrices[i] [ch, f_{i+1}, f_i], biases[i] [ch, f_{i+1}, 1], factors[i] [ch, f_{i+1}, 1].
    """
    N, ch = x.shape
    pidx = torch.from_numpy(entropy_param_index(0, ch, N))  # [2, ch, N]
    xt = x.t()  # [ch, N]
    inp = torch.stack([xt - 0.5 * q, xt + 0.5 * q], 0)  # [2, ch, N]
    h = inp[..., None]  # [2,ch,N,1]
    for i in range(len(matrices)):
        Wm = F.softplus(matrices[i])[pidx]  # [2,ch,N,fo,fi]
        h = (Wm @ h[..., None]).squeeze(-1) + biases[i][pidx].squeeze(-1)
        if i < len(factors):
            h = h + torch.tanh(factors[i][pidx].squeeze(-1)) * torch.tanh(h)
    logits = h.squeeze(-1)  # [2,ch,N]
    lower, upper = logits[0], logits[1]
    sign = -(lower + upper).sign().detach()
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    lik = _LowerBound.apply(lik, likelihood_bound)
    return (-torch.log2(lik)).t()


# --------------------------------------------------------------------------------------
# photometric loss: L1 + SSIM, examples/simple_trainer.py:948-953
# --------------------------------------------------------------------------------------
# PARITY UNPINNED: `fused_ssim` is a third-party dependency that is absent from /root/reference
# (examples/requirements.txt:26 pins rahul-goel/fused-ssim@1272e21a) and the reference holds no test or golden
# vector for it.  Restated from its published algorithm: SSIM (Wang et al. 2004) with an 11-tap Gaussian window of
# sigma 1.5 applied separably with ZERO padding, C1 = 0.01^2, C2 = 0.03^2, per channel; padding="valid" drops the
# 5-pixel border of the map before the mean; only the first image receives a gradient.  tests/test_oracle_golden.py
# checks this restatement against a brute-force float64 evaluation of the definition and by finite differences.
def ssim_window(dtype=torch.float64) -> Tensor:
    d = torch.arange(11, dtype=torch.float64) - 5
    g = torch.exp(-(d * d) / (2 * 1.5 * 1.5))
    return (g / g.sum()).to(dtype)


def ssim_map(img1: Tensor, img2: Tensor) -> Tensor:
    """SSIM map [B,CH,H,W] of [B,CH,H,W] images (zero-padded Gaussian window, "same" size)."""
    B, CH, H, W = img1.shape
    w1 = ssim_window(img1.dtype)
    win = (w1[:, None] * w1[None, :]).expand(CH, 1, 11, 11).contiguous()

    def blur(t):
        return F.conv2d(t, win, padding=5, groups=CH)

    mu1, mu2 = blur(img1), blur(img2)
    s1 = blur(img1 * img1) - mu1 * mu1
    s2 = blur(img2 * img2) - mu2 * mu2
    s12 = blur(img1 * img2) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))


def fused_ssim(img1: Tensor, img2: Tensor, padding: str = "same") -> Tensor:
    """Mean SSIM of [B,CH,H,W] images; the second image is treated as a constant."""
    m = ssim_map(img1, img2.detach())
    if padding == "valid":
        m = m[:, :, 5:-5, 5:-5]
    return m.mean()


def photometric_loss(colors: Tensor, pixels: Tensor, ssim_lambda: float = 0.2, padding: str = "valid") -> Tensor:
    """The trainer's loss on [B,H,W,CH] images (examples/simple_trainer.py:948-953)."""
    l1 = (colors - pixels).abs().mean()
    ssim = fused_ssim(colors.permute(0, 3, 1, 2), pixels.permute(0, 3, 1, 2), padding)
    return l1 * (1.0 - ssim_lambda) + (1.0 - ssim) * ssim_lambda


# --------------------------------------------------------------------------------------
# optimizer step: gsplat/cuda/csrc/adam.cu:13-43
# --------------------------------------------------------------------------------------
def selective_adam_update(param, grad, exp_avg, exp_avg_sq, visible, lr, b1, b2, eps):
    """In-place reference update of the rows flagged in `visible` [N] (no bias correction, adam.cu:31-40)."""
    sel = visible.view(-1, *([1] * (param.dim() - 1))).expand_as(param)
    m = b1 * exp_avg + (1.0 - b1) * grad
    v = b2 * exp_avg_sq + (1.0 - b2) * grad * grad
    step = -lr * m / (v.sqrt() + eps)
    param.copy_(torch.where(sel, param + step, param))
    exp_avg.copy_(torch.where(sel, m, exp_avg))
    exp_avg_sq.copy_(torch.where(sel, v, exp_avg_sq))


# --------------------------------------------------------------------------------------
# un-fused geometry ops, pinned against the reference's _torch_impl.py:41-247 by
# tests/golden/ref_cpu_geom_ops.npz (generator: oracle/make_golden_geom.py)
# --------------------------------------------------------------------------------------
def _triu6(m: Tensor) -> Tensor:
    """[...,3,3] -> [...,6]: the mean of the mirrored entries (_torch_impl.py:55-59)."""
    iu, ju = [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]
    return 0.5 * (m[..., iu, ju] + m[..., ju, iu])


def quat_scale_to_covar_preci(quats: Tensor, scales: Tensor, compute_covar=True, compute_preci=True, triu=False):
    R = quat_to_rotmat(quats)
    covars = precis = None
    if compute_covar:
        M = R * scales[..., None, :]
        covars = M @ M.transpose(-1, -2)
        covars = _triu6(covars) if triu else covars
    if compute_preci:
        P = R / scales[..., None, :]
        precis = P @ P.transpose(-1, -2)
        precis = _triu6(precis) if triu else precis
    return covars, precis


def proj(means: Tensor, covars: Tensor, Ks: Tensor, width: int, height: int, camera_model: str = "pinhole"):
    """Camera-space means [C,N,3] / full covars [C,N,3,3] -> means2d [C,N,2], covars2d [C,N,2,2] = J covar J^T
    (proj.cuh:9-36 ortho, :80-119 pinhole with the fov clamp of the Jacobian, :202-241 fisheye)."""
    C, N, _ = means.shape
    x, y, z = means.unbind(-1)
    fx, fy = Ks[:, 0, 0, None], Ks[:, 1, 1, None]
    cx, cy = Ks[:, 0, 2, None], Ks[:, 1, 2, None]
    O = torch.zeros_like(z)
    if camera_model == "pinhole":
        tan_fovx, tan_fovy = 0.5 * width / fx, 0.5 * height / fy
        lim_xp, lim_xn = (width - cx) / fx + 0.3 * tan_fovx, cx / fx + 0.3 * tan_fovx
        lim_yp, lim_yn = (height - cy) / fy + 0.3 * tan_fovy, cy / fy + 0.3 * tan_fovy
        rz = 1.0 / z
        tx = z * torch.minimum(lim_xp.expand_as(z), torch.maximum(-lim_xn.expand_as(z), x * rz))
        ty = z * torch.minimum(lim_yp.expand_as(z), torch.maximum(-lim_yn.expand_as(z), y * rz))
        J = torch.stack([fx * rz, O, -fx * tx * rz * rz, O, fy * rz, -fy * ty * rz * rz], -1)
        means2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], -1)
    elif camera_model == "ortho":
        J = torch.stack([fx + O, O, O, O, fy + O, O], -1)
        means2d = torch.stack([fx * x + cx, fy * y + cy], -1)
    elif camera_model == "fisheye":
        eps = 1e-7
        xy_len = torch.sqrt(x * x + y * y) + eps
        theta = torch.atan2(xy_len, z + eps)
        means2d = torch.stack([x * fx * theta / xy_len + cx, y * fy * theta / xy_len + cy], -1)
        x2, y2, xy = x * x + eps, y * y, x * y
        x2y2 = x2 + y2
        inv = 1.0 / (x2y2 + z * z)
        b = torch.atan2(xy_len, z) / xy_len / x2y2
        a = z * inv / x2y2
        J = torch.stack([fx * (x2 * a + y2 * b), fx * xy * (a - b), -fx * x * inv,
                         fy * xy * (a - b), fy * (y2 * a + x2 * b), -fy * y * inv], -1)
    else:
        raise ValueError(camera_model)
    J = J.reshape(C, N, 2, 3)
    return means2d, J @ covars @ J.transpose(-1, -2)


def world_to_cam(means: Tensor, covars: Tensor, viewmats: Tensor):
    """means_c = R p + t, covars_c = R covar R^T for every camera (transform.cuh:9-49)."""
    R, t = viewmats[:, :3, :3], viewmats[:, :3, 3]
    means_c = torch.einsum("cij,nj->cni", R, means) + t[:, None, :]
    covars_c = torch.einsum("cij,njk,clk->cnil", R, covars, R)
    return means_c, covars_c


def compute_relocation(opacities: Tensor, scales: Tensor, ratios: Tensor, binoms: Tensor):
    """gsplat/cuda/csrc/compute_relocation.cu:13-40 restated (float64 inside); ratios already clamped to [1, n_max]."""
    n_max = binoms.shape[0]
    o = opacities.double()
    n = ratios.long()
    new_o = 1.0 - (1.0 - o) ** (1.0 / n.double())
    k = torch.arange(n_max, dtype=torch.float64)
    term = ((-1.0) ** k) / torch.sqrt(k + 1.0) * new_o[:, None] ** (k + 1.0)  # [N, n_max]
    row_sums = torch.einsum("ik,nk->ni", torch.tril(binoms.double()), term)  # sum_k binom(i-1, k) term_k, k <= i-1
    use = (torch.arange(n_max)[None, :] < n[:, None]).double()  # rows i-1 = 0 .. n-1
    denom = (row_sums * use).sum(-1)
    return new_o.to(opacities.dtype), (scales.double() * (o / denom)[:, None]).to(scales.dtype)


# --------------------------------------------------------------------------------------
# dynamic splats: temporal pre-transform, examples/simple_trainer_dyngs.py:476-520 restated
# --------------------------------------------------------------------------------------
def stg_temporal_transform(means, quats, scales, opacities, trbf_center, trbf_scale, motion, omega, timestamp: float):
    d = timestamp - trbf_center.reshape(-1, 1)  # [N,1]
    x = d / (math.sqrt(2) * torch.exp(trbf_scale.reshape(-1, 1)))
    trbf = torch.exp(-x * x)
    opacity = torch.sigmoid(opacities) * trbf.squeeze(-1)
    dd = d.detach()
    means_t = means + motion[:, 0:3] * dd + motion[:, 3:6] * dd * dd + motion[:, 6:9] * dd * dd * dd
    quats_t = F.normalize(quats + dd * omega)
    return means_t, quats_t, torch.exp(scales), opacity, trbf.squeeze(-1).detach()


# --------------------------------------------------------------------------------------
# rasterize_to_indices_in_range (gsplat/cuda/csrc/rasterize_to_indices_in_range.cu:13-176) restated; python loops,
# for SMALL cases only
# --------------------------------------------------------------------------------------
def rasterize_to_indices_in_range(range_start, range_end, transmittances, means2d, conics, opacities, width, height,
                                  tile_size, isect_offsets, flatten_ids):
    """-> (gaussian_ids, pixel_ids, camera_ids) int64, per pixel (camera-major, row-major) front to back.  A slice is
    counted in batches of tile_size^2 intersections per tile; float32 arithmetic like the kernel (expf, not __expf)."""
    C, N = means2d.shape[0], means2d.shape[1]
    th, tw = isect_offsets.shape[1:3]
    m2 = means2d.reshape(-1, 2).numpy().astype(np.float32)
    cn = conics.reshape(-1, 3).numpy().astype(np.float32)
    op = opacities.reshape(-1).numpy().astype(np.float32)
    offs = isect_offsets.reshape(-1).numpy()
    fl = flatten_ids.numpy()
    trans_in = transmittances.numpy().astype(np.float32)
    n_isects = fl.shape[0]
    block = tile_size * tile_size
    gids, pids, cids = [], [], []
    f32 = np.float32
    for c in range(C):
        for i in range(height):
            for j in range(width):
                tile = (i // tile_size) * tw + (j // tile_size)
                ct = c * th * tw + tile
                lo = int(offs[ct])
                hi = n_isects if ct == C * th * tw - 1 else int(offs[ct + 1])
                nb = (hi - lo + block - 1) // block
                if range_start >= nb:
                    continue
                a, b = lo + block * range_start, min(hi, lo + block * min(range_end, nb))
                T = trans_in[c, i, j]
                px, py = f32(j + 0.5), f32(i + 0.5)
                for idx in range(a, b):
                    g = int(fl[idx])
                    dx, dy = m2[g, 0] - px, m2[g, 1] - py
                    sigma = f32(0.5) * (cn[g, 0] * dx * dx + cn[g, 2] * dy * dy) + cn[g, 1] * dx * dy
                    alpha = min(f32(0.999), op[g] * np.exp(-sigma, dtype=np.float32))
                    if sigma < 0 or alpha < f32(1.0 / 255.0):
                        continue
                    nT = T * (f32(1.0) - alpha)
                    if nT <= f32(1e-4):
                        break
                    gids.append(g % N)
                    pids.append(i * width + j)
                    cids.append(c)
                    T = nT
    mk = lambda v: torch.tensor(v, dtype=torch.int64)
    return mk(gids), mk(pids), mk(cids)
