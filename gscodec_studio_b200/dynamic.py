"""Temporal pre-transform of Spacetime-Gaussian splats (BASELINE config 4): the
per-timestamp attribute chain the dynamic trainer runs before ``rasterization`` --
examples/simple_trainer_dyngs.py:476-520 -- as one forward and one backward kernel (additive op)."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from ._cabi import aligned, call, device_guard, ptr, stream_ptr


def stg_temporal_transform(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]
    scales: Tensor,  # [N, 3]  log-scales (the trainer's parameter)
    opacities: Tensor,  # [N]  logits
    trbf_center: Tensor,  # [N, 1]
    trbf_scale: Tensor,  # [N, 1]  log of the temporal radius
    motion: Tensor,  # [N, 9]
    omega: Tensor,  # [N, 4]
    timestamp: float,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (means_t [N,3], quats_t [N,4] normalised, scales [N,3] = exp, opacities [N] = sigmoid * trbf, trbf [N]).

    ``trbf`` (the temporal radial basis value, used for the trainer's ``trbf > 0.05`` visibility mask) is returned
    without a graph; the other four outputs are differentiable w.r.t. all eight inputs exactly like the torch
    chain of the reference (the time offset inside the motion polynomial and the rotation is detached there)."""
    N = means.shape[0]
    assert means.shape == (N, 3) and quats.shape == (N, 4) and scales.shape == (N, 3), (means.shape, quats.shape, scales.shape)
    assert opacities.shape == (N,), opacities.shape
    assert trbf_center.numel() == N and trbf_scale.numel() == N, (trbf_center.shape, trbf_scale.shape)
    assert motion.shape == (N, 9) and omega.shape == (N, 4), (motion.shape, omega.shape)
    return _StgTemporal.apply(means, quats, scales, opacities, trbf_center, trbf_scale, motion, omega, float(timestamp))


def _f32(t: Tensor) -> Tensor:
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError("stg_temporal_transform: float32 CUDA tensors expected (no CPU implementation)")
    return aligned(t)


class _StgTemporal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, opacities, trbf_center, trbf_scale, motion, omega, timestamp):
        means, quats, scales, opacities, trbf_center, trbf_scale, motion, omega = map(
            _f32, (means, quats, scales, opacities, trbf_center, trbf_scale, motion, omega))
        N, dev = means.shape[0], means.device
        means_t, quats_t, scales_o = torch.empty_like(means), torch.empty_like(quats), torch.empty_like(scales)
        opac_o, trbf = torch.empty_like(opacities), torch.empty_like(opacities)
        with device_guard(dev):
            call("gsb200_stg_temporal_fwd", N, timestamp, ptr(means), ptr(quats), ptr(scales), ptr(opacities), ptr(trbf_center),
                 ptr(trbf_scale), ptr(motion), ptr(omega), ptr(means_t), ptr(quats_t), ptr(scales_o), ptr(opac_o), ptr(trbf),
                 stream_ptr(dev))
        ctx.save_for_backward(quats, scales, opacities, trbf_center, trbf_scale, omega)
        ctx.timestamp = timestamp
        ctx.mark_non_differentiable(trbf)
        return means_t, quats_t, scales_o, opac_o, trbf

    @staticmethod
    def backward(ctx, v_means_t, v_quats_t, v_scales, v_opacities, _v_trbf):
        quats, scales, opacities, trbf_center, trbf_scale, omega = ctx.saved_tensors
        N, dev = quats.shape[0], quats.device
        v_means_t, v_quats_t, v_scales, v_opacities = (_f32(t.contiguous()) for t in (v_means_t, v_quats_t, v_scales, v_opacities))
        v_means, v_quats, v_scales_raw = torch.empty_like(v_means_t), torch.empty_like(quats), torch.empty_like(scales)
        v_opac_raw, v_center, v_tscale = torch.empty_like(opacities), torch.empty_like(trbf_center), torch.empty_like(trbf_scale)
        v_motion = torch.empty((N, 9), dtype=torch.float32, device=dev)
        v_omega = torch.empty_like(omega)
        with device_guard(dev):
            call("gsb200_stg_temporal_bwd", N, ctx.timestamp, ptr(quats), ptr(scales), ptr(opacities), ptr(trbf_center),
                 ptr(trbf_scale), ptr(omega), ptr(v_means_t), ptr(v_quats_t), ptr(v_scales), ptr(v_opacities), ptr(v_means),
                 ptr(v_quats), ptr(v_scales_raw), ptr(v_opac_raw), ptr(v_center), ptr(v_tscale), ptr(v_motion), ptr(v_omega),
                 stream_ptr(dev))
        return v_means, v_quats, v_scales_raw, v_opac_raw, v_center, v_tscale, v_motion, v_omega, None
