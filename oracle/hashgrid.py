"""CPU oracle of the hash-grid ("gaussian_model") entropy model.  TEST INFRASTRUCTURE ONLY -- preparation for the
SURVEY.md section 8(f).3 row; nothing under ``gscodec_studio_b200/`` imports it and no CUDA counterpart exists yet.

PARITY UNPINNED: the encoder is the reference's third-party CUDA extension ``_gridencoder``
(/root/reference/third_party/gridencoder/src/gridencoder.cu, imported by
gsplat/compression_simulation/gaussian_distribution_model.py:8), which has no CPU implementation, so nothing here
could be checked against the reference in this container.  ``oracle/Makefile`` (target ``ref_grid``) compiles that
extension unmodified into ``oracle/_ref/_gridencoder.so``; ``check_against_reference()`` below is the pinning run
to execute on a GPU box.

Restated (written from the reference's behaviour, not copied):
* grid layout                GridEncoder.__init__, gaussian_distribution_model.py:204-233
* index / hash               get_grid_index + fast_hash, gridencoder.cu:45-88
* forward interpolation      kernel_grid, gridencoder.cu:100-358 (binary_vxl = None, PV = 0, min_level_id = 0: the only
                             way compression_simulation calls it): positions scaled to ``x (R - 2) + 0.5``, the 2^D
                             corners, corners ON the outer shell (coordinate 0 or R-1) dropped and the remaining
                             weights renormalised, out-of-range inputs -> 0
* input Jacobian             the dy_dx block of kernel_grid, gridencoder.cu:360-660: plain multilinear derivative with
                             factor (R - 2), dropped corners read as 0, NOT the derivative of the renormalised forward
* embedding gradient         kernel_grid_backward, gridencoder.cu:664-853 (scatter-add of w / wn)
* STE_binary, mix_3D2D_encoding, hash_based_estimator, Entropy_gaussian
                             gaussian_distribution_model.py:19-39, 276-382; entropy_model.py:313-345
"""
from __future__ import annotations

import itertools
import math
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

_PRIMES = (1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737)
_U32 = (1 << 32) - 1


def grid_layout(num_dim: int, resolutions: Sequence[int], log2_hashmap_size: int) -> List[int]:
    """Offsets [n_levels + 1] of the per-level tables: min(2^log2, R^D) entries rounded up to a multiple of 8."""
    offsets, offset = [], 0
    for r in resolutions:
        n = min(2 ** log2_hashmap_size, int(r) ** num_dim)
        n = int(math.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return offsets


def grid_index(pos_grid: Tensor, hashmap_size: int, resolution: int) -> Tensor:
    """Table row of integer grid coordinates [..., D] (int64 in, uint32 arithmetic): dense row-major index while the
    running stride fits the table, the XOR-of-primes hash otherwise; always modulo the table size."""
    D = pos_grid.shape[-1]
    stride, index = 1, torch.zeros(pos_grid.shape[:-1], dtype=torch.int64)
    for d in range(D):
        if stride > hashmap_size:
            break
        index = (index + pos_grid[..., d] * stride) & _U32
        stride = (stride * resolution) & _U32
    if stride > hashmap_size:
        index = torch.zeros_like(index)
        for d in range(D):
            index = index ^ ((pos_grid[..., d] * _PRIMES[d]) & _U32)
    return index % hashmap_size


def _level_terms(x: Tensor, resolution: int):
    """x [N, D] float32 in [0,1] -> integer cell origin [N,D], fractional position [N,D] (float32 like the kernel)."""
    pos = x.float() * float(resolution - 2) + 0.5
    cell = torch.floor(pos)
    return cell.long(), (pos - cell)


def encode_level(x: Tensor, table: Tensor, resolution: int, want_jacobian: bool):
    """One level.  x [N,D], table [S,F] -> out [N,F], (rows [N,2^D], weights/wn [N,2^D] for the backward),
    dy_dx [N,D,F] or None."""
    N, D = x.shape
    S, F = table.shape
    oob = ((x < 0) | (x > 1)).any(-1)
    cell, frac = _level_terms(x, resolution)
    corners = list(itertools.product((0, 1), repeat=D))  # corner idx bit d = corners[.][d]; idx = sum bit_d << d
    rows = torch.zeros(N, 1 << D, dtype=torch.int64)
    wts = torch.zeros(N, 1 << D, dtype=torch.float32)
    valid = torch.zeros(N, 1 << D, dtype=torch.bool)
    for bits in corners:
        idx = sum(b << d for d, b in enumerate(bits))
        w = torch.ones(N, dtype=torch.float32)
        pg = torch.empty(N, D, dtype=torch.int64)
        for d, b in enumerate(bits):
            w = w * (frac[:, d] if b else (1 - frac[:, d]))
            pg[:, d] = torch.clamp(cell[:, d] + 1, max=resolution - 1) if b else cell[:, d]
        shell = ((pg == 0) | (pg == resolution - 1)).any(-1)
        rows[:, idx] = grid_index(pg, S, resolution)
        wts[:, idx] = w
        valid[:, idx] = ~shell
    wn = (wts * valid).sum(-1)
    wn = torch.where(wn == 0, wn + 1e-9, wn)
    coef = torch.where(valid, wts * (1.0 / wn)[:, None], torch.zeros_like(wts))  # w * wn_re, dropped corners 0
    coef = torch.where(oob[:, None], torch.zeros_like(coef), coef)
    out = (coef[..., None] * table[rows]).sum(1)  # [N, F]
    jac = None
    if want_jacobian:
        jac = torch.zeros(N, D, F, dtype=torch.float32)
        for gd in range(D):
            others = [d for d in range(D) if d != gd]
            for bits in itertools.product((0, 1), repeat=D - 1):
                w = torch.full((N,), float(resolution - 2), dtype=torch.float32)
                pg = torch.empty(N, D, dtype=torch.int64)
                for d, b in zip(others, bits):
                    w = w * (frac[:, d] if b else (1 - frac[:, d]))
                    pg[:, d] = torch.clamp(cell[:, d] + 1, max=resolution - 1) if b else cell[:, d]
                vals = []
                for side in (0, 1):
                    pg[:, gd] = torch.clamp(cell[:, gd] + 1, max=resolution - 1) if side else cell[:, gd]
                    shell = ((pg == 0) | (pg == resolution - 1)).any(-1)
                    v = table[grid_index(pg, S, resolution)]
                    vals.append(torch.where(shell[:, None], torch.zeros_like(v), v))
                jac[:, gd] += w[:, None] * (vals[1] - vals[0])
        jac = torch.where(oob[:, None, None], torch.zeros_like(jac), jac)
    return out, rows, coef, jac


class _GridEncode(torch.autograd.Function):
    """grid_encode(inputs [N,D], embeddings [S_total,F]) -> [N, L*F] (level-major), with the reference's gradients."""

    @staticmethod
    def forward(ctx, inputs: Tensor, embeddings: Tensor, offsets: Tuple[int, ...], resolutions: Tuple[int, ...]):
        want_jac = inputs.requires_grad
        outs, saved = [], []
        for lvl, res in enumerate(resolutions):
            table = embeddings[offsets[lvl]:offsets[lvl + 1]]
            out, rows, coef, jac = encode_level(inputs.detach(), table.detach(), int(res), want_jac)
            outs.append(out)
            saved.append((rows + offsets[lvl], coef, jac))
        ctx.saved = saved
        ctx.shape = tuple(embeddings.shape)
        return torch.cat(outs, dim=1)

    @staticmethod
    def backward(ctx, grad: Tensor):
        F = ctx.shape[1]
        g_emb = torch.zeros(ctx.shape, dtype=grad.dtype)
        g_in = None
        for lvl, (rows, coef, jac) in enumerate(ctx.saved):
            g = grad[:, lvl * F:(lvl + 1) * F]  # [N, F]
            contrib = coef[..., None].to(grad.dtype) * g[:, None, :]  # [N, 2^D, F]
            g_emb.index_add_(0, rows.reshape(-1), contrib.reshape(-1, F))
            if jac is not None:
                part = (jac.to(grad.dtype) * g[:, None, :]).sum(-1)  # [N, D]
                g_in = part if g_in is None else g_in + part
        return g_in, g_emb, None, None


def grid_encode(inputs: Tensor, embeddings: Tensor, offsets: Sequence[int], resolutions: Sequence[int]) -> Tensor:
    return _GridEncode.apply(inputs, embeddings, tuple(int(o) for o in offsets), tuple(int(r) for r in resolutions))


class _SteBinary(torch.autograd.Function):
    """sign with +1 at 0, gradient passed where |x| <= 1 (gaussian_distribution_model.py:19-38)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.where(torch.clamp(x, -1, 1) >= 0, torch.ones_like(x), -torch.ones_like(x))

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * ((x >= -1) & (x <= 1)).to(g.dtype)


class GridEncoder(nn.Module):
    def __init__(self, num_dim: int, n_features: int, resolutions: Sequence[int], log2_hashmap_size: int, ste_binary: bool = True):
        super().__init__()
        self.num_dim, self.n_features, self.ste_binary = num_dim, n_features, ste_binary
        self.resolutions = [int(r) for r in resolutions]
        self.offsets = grid_layout(num_dim, self.resolutions, log2_hashmap_size)
        self.output_dim = len(self.resolutions) * n_features
        self.params = nn.Parameter(torch.empty(self.offsets[-1], n_features).uniform_(-1e-4, 1e-4))

    def forward(self, x: Tensor) -> Tensor:
        emb = _SteBinary.apply(self.params) if self.ste_binary else self.params
        return grid_encode(x.reshape(-1, self.num_dim), emb, self.offsets, self.resolutions)


class HashBasedEstimator(nn.Module):
    """hash_based_estimator (gaussian_distribution_model.py:348-382): 3D grid + three 2D planes -> MLP -> (mean, scale)."""

    def __init__(self, channel: int):
        super().__init__()
        self.channel = channel
        r3 = (18, 24, 33, 44, 59, 80, 108, 148, 201, 275, 376, 514)
        r2 = (130, 258, 514, 1026)
        self.encoding_xyz = GridEncoder(3, 2, r3, 13)
        self.encoding_xy = GridEncoder(2, 2, r2, 15)
        self.encoding_xz = GridEncoder(2, 2, r2, 15)
        self.encoding_yz = GridEncoder(2, 2, r2, 15)
        dim = self.encoding_xyz.output_dim + 3 * self.encoding_xy.output_dim
        self.mlp_regressor = nn.Sequential(nn.Linear(dim, channel * 5), nn.ReLU(True), nn.Linear(channel * 5, channel * 2))

    def forward(self, pos: Tensor):
        x = (pos + 100.0) / 200.0  # bound_min = -100, bound_max = 100
        xx, yy, zz = x[:, 0:1], x[:, 1:2], x[:, 2:3]
        feats = torch.cat([self.encoding_xyz(x), self.encoding_xy(torch.cat([xx, yy], -1)),
                           self.encoding_xz(torch.cat([xx, zz], -1)), self.encoding_yz(torch.cat([yy, zz], -1))], -1)
        out = self.mlp_regressor(feats)
        return out[:, :self.channel], out[:, self.channel:2 * self.channel]


def entropy_gaussian_bits(x: Tensor, q: float, mean: Tensor, scale: Tensor, likelihood_bound: float = 1e-6) -> Tensor:
    """Entropy_gaussian.forward (entropy_model.py:322-339): -log2 of the Normal(mean, max(scale, 1e-9)) mass of
    [x - q/2, x + q/2], lower-bounded (the bound's straight-through gradient is the factorized model's LowerBound)."""
    from oracle.oracle import _LowerBound

    scale = torch.clamp(scale, min=1e-9)
    normal = torch.distributions.normal.Normal(mean, scale)
    like = torch.abs(normal.cdf(x + 0.5 * q) - normal.cdf(x - 0.5 * q))
    return -torch.log2(_LowerBound.apply(like, likelihood_bound))


def check_against_reference(n: int = 4096, seed: int = 0) -> None:  # pragma: no cover - needs a GPU box
    """The pinning run (GPU box, after `make -C oracle ref_grid`): forward, embedding gradient and input gradient of
    the 3D and 2D encoders against the unmodified reference extension on random positions incl. shell / out-of-range."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))
    import _gridencoder as ref  # noqa

    g = torch.Generator().manual_seed(seed)
    for D, res, log2 in ((3, (18, 24, 33, 44, 59, 80, 108, 148, 201, 275, 376, 514), 13), (2, (130, 258, 514, 1026), 15)):
        offs = grid_layout(D, res, log2)
        x = torch.rand(n, D, generator=g)
        x[:8] = torch.tensor([0.0, 1.0, 1.2, -0.1, 0.5, 0.999, 0.001, 0.25])[:, None]
        emb = torch.randn(offs[-1], 2, generator=g)
        xi, ei = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
        out = grid_encode(xi, ei, offs, res)
        w = torch.randn(out.shape, generator=g)
        (out * w).sum().backward()
        L, F = len(res), 2
        dev = torch.device("cuda")
        o_ref = torch.empty(L, n, F, device=dev)
        dy_dx = torch.empty(n, L * D * F, device=dev)
        offs_t, res_t = torch.tensor(offs, dtype=torch.int32, device=dev), torch.tensor(res, dtype=torch.int32, device=dev)
        ref.grid_encode_forward(x.to(dev), emb.to(dev), offs_t, res_t, o_ref, n, D, F, L, 0, 128, 0.0, dy_dx, None, None)
        o_ref2 = o_ref.permute(1, 0, 2).reshape(n, L * F).cpu()
        g_emb, g_in = torch.zeros_like(emb, device=dev), torch.zeros(n, D, device=dev)
        grad = w.view(n, L, F).permute(1, 0, 2).contiguous().to(dev)
        ref.grid_encode_backward(grad, x.to(dev), emb.to(dev), offs_t, res_t, g_emb, n, D, F, L, 0, 128, dy_dx, g_in, None, None)
        for name, a, b in (("forward", out.detach(), o_ref2), ("grad embeddings", ei.grad, g_emb.cpu()), ("grad inputs", xi.grad, g_in.cpu())):
            err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)
            print(f"D={D} {name}: rel err {err:.2e}")
            assert err < 1e-4, (D, name, err)
    print("hash-grid oracle pinned against the reference extension")


if __name__ == "__main__":  # pragma: no cover
    check_against_reference()
