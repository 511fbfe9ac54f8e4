"""Synthetic scene builder used by tests and bench.py.

Same semantics as the reference's ``gsplat/_helper.py::load_test_data`` (:10-55): crop the garden point
cloud to an axis-aligned box, replicate it on a ``scene_grid x scene_grid`` lattice in x/y, draw random
scales U(0, 0.02), normalised N(0,1) quaternions and U(0,1) opacities.  Unlike the reference the random
draws take an explicit seed and happen on the CPU generator, so CPU oracle and GPU runs see identical
inputs on every machine.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch

ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "test_garden.npz")


def load_test_data(
    data_path: Optional[str] = None,
    device="cuda",
    scene_crop: Tuple[float, float, float, float, float, float] = (-2, -2, -2, 2, 2, 2),
    scene_grid: int = 1,
    seed: int = 42,
):
    assert scene_grid % 2 == 1, "scene_grid must be odd"
    data = np.load(data_path or ASSET)
    height, width = int(data["height"]), int(data["width"])
    viewmats = torch.from_numpy(data["viewmats"]).float()
    Ks = torch.from_numpy(data["Ks"]).float()
    pts = torch.from_numpy(data["means3d"]).float()
    rgb = torch.from_numpy(data["colors"] / 255.0).float()

    lo = torch.tensor(scene_crop[:3], dtype=torch.float32)
    hi = torch.tensor(scene_crop[3:], dtype=torch.float32)
    keep = ((pts >= lo) & (pts <= hi)).all(dim=-1)
    pts, rgb = pts[keep], rgb[keep]

    half = scene_grid // 2
    steps = torch.arange(-half, half + 1, dtype=torch.float32)
    gx, gy = torch.meshgrid(steps, steps, indexing="ij")
    shifts = torch.stack([gx, gy, torch.zeros_like(gx)], dim=-1).reshape(-1, 3) * (hi - lo)
    means = (pts[None] + shifts[:, None]).reshape(-1, 3)
    colors = rgb.repeat(scene_grid * scene_grid, 1)

    n = means.shape[0]
    gen = torch.Generator().manual_seed(seed)
    scales = torch.rand((n, 3), generator=gen) * 0.02
    quats = torch.nn.functional.normalize(torch.randn((n, 4), generator=gen), dim=-1)
    opacities = torch.rand((n,), generator=gen)
    to = lambda t: t.to(device).contiguous()
    return to(means), to(quats), to(scales), to(opacities), to(colors), to(viewmats), to(Ks), width, height
