"""Small torch helpers of the reference's ``gsplat/utils.py`` that the 2DGS render path needs (depth -> points -> normals,
:44-131) plus the quaternion / projection-matrix helpers callers import from there (:8-41, :134-155)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import Tensor


def normalized_quat_to_rotmat(quat: Tensor) -> Tensor:
    """Normalized wxyz quaternion (..., 4) -> rotation matrix (..., 3, 3)."""
    assert quat.shape[-1] == 4, quat.shape
    w, x, y, z = torch.unbind(quat, dim=-1)
    rows = [1 - 2 * (y**2 + z**2), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x**2 + z**2), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x**2 + y**2)]
    return torch.stack(rows, dim=-1).reshape(quat.shape[:-1] + (3, 3))


def log_transform(x):
    return torch.sign(x) * torch.log1p(torch.abs(x))


def inverse_log_transform(y):
    return torch.sign(y) * (torch.expm1(torch.abs(y)))


def depth_to_points(depths: Tensor, camtoworlds: Tensor, Ks: Tensor, z_depth: bool = True) -> Tensor:
    """Depth maps [..., H, W, 1] -> world-space points [..., H, W, 3] (pixel centres at +0.5; `z_depth` False: the depth
    is measured along the ray)."""
    assert depths.shape[-1] == 1, f"Invalid depth shape: {depths.shape}"
    assert camtoworlds.shape[-2:] == (4, 4), f"Invalid viewmats shape: {camtoworlds.shape}"
    assert Ks.shape[-2:] == (3, 3), f"Invalid Ks shape: {Ks.shape}"
    assert depths.shape[:-3] == camtoworlds.shape[:-2] == Ks.shape[:-2], \
        f"Shape mismatch! depths: {depths.shape}, viewmats: {camtoworlds.shape}, Ks: {Ks.shape}"
    height, width = depths.shape[-3:-1]
    x, y = torch.meshgrid(torch.arange(width, device=depths.device), torch.arange(height, device=depths.device), indexing="xy")
    fx, fy, cx, cy = Ks[..., 0, 0], Ks[..., 1, 1], Ks[..., 0, 2], Ks[..., 1, 2]
    dirs_cam = torch.stack([(x - cx[..., None, None] + 0.5) / fx[..., None, None],
                            (y - cy[..., None, None] + 0.5) / fy[..., None, None]], dim=-1)
    dirs_cam = F.pad(dirs_cam, (0, 1), value=1.0)  # [..., H, W, 3]
    directions = torch.einsum("...ij,...hwj->...hwi", camtoworlds[..., :3, :3], dirs_cam)
    if not z_depth:
        directions = F.normalize(directions, dim=-1)
    return camtoworlds[..., :3, -1][..., None, None, :] + depths * directions


def depth_to_normal(depths: Tensor, camtoworlds: Tensor, Ks: Tensor, z_depth: bool = True) -> Tensor:
    """Depth maps [..., H, W, 1] -> world-space surface normals [..., H, W, 3] from central differences of the
    back-projected points (zero on the one-pixel border)."""
    points = depth_to_points(depths, camtoworlds, Ks, z_depth=z_depth)
    dx = points[..., 2:, 1:-1, :] - points[..., :-2, 1:-1, :]
    dy = points[..., 1:-1, 2:, :] - points[..., 1:-1, :-2, :]
    normals = F.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    return F.pad(normals, (0, 0, 1, 1, 1, 1), value=0.0)


def get_projection_matrix(znear, zfar, fovX, fovY, device="cuda"):
    """OpenGL-style projection matrix (used by the Inria-backend wrappers of the reference)."""
    top, right = math.tan(fovY / 2) * znear, math.tan(fovX / 2) * znear
    P = torch.zeros(4, 4, device=device)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P
