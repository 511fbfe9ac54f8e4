"""Photometric loss of the training step, fused.

Reference call site (examples/simple_trainer.py:948-953; same in simple_trainer_dyngs.py:645-651 and
simple_trainer_STG.py:692-698)::

    l1loss   = F.l1_loss(colors, pixels)
    ssimloss = 1.0 - fused_ssim(colors.permute(0, 3, 1, 2), pixels.permute(0, 3, 1, 2), padding="valid")
    loss     = l1loss * (1.0 - cfg.ssim_lambda) + ssimloss * cfg.ssim_lambda

``fused_ssim`` there is the un-vendored third-party package pinned in examples/requirements.txt:26
(rahul-goel/fused-ssim@1272e21a).  This module provides

* ``fused_ssim(img1, img2, padding="same", train=True)`` -- same signature and semantics as that package
  ([B,CH,H,W] inputs, mean SSIM, gradient to ``img1`` only), so the trainers' import can point here;
* ``photometric_loss(colors, pixels, ssim_lambda)`` -- the whole three-line chain above as one forward and one
  backward kernel on the renderer's [B,H,W,CH] layout (additive).

Both run the CUDA kernels of ``csrc/loss.cu`` through the C-ABI; there is no CPU implementation.
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _cabi
from ._cabi import call, device_guard, ptr, stream_ptr


def _check(img: Tensor, name: str) -> Tensor:
    if not img.is_cuda:
        raise RuntimeError(f"gscodec_studio_b200: {name} must be a CUDA tensor (the loss has no CPU implementation)")
    if img.dtype != torch.float32:
        raise RuntimeError(f"gscodec_studio_b200: expected float32 {name}, got {img.dtype}")
    return img.contiguous()


class _Photometric(torch.autograd.Function):
    """out = coef_l1 * mean|img1-img2| + coef_ssim * mean(ssim) + bias on [B,H,W,CH] images."""

    @staticmethod
    def forward(ctx, img1: Tensor, img2: Tensor, coef_l1: float, coef_ssim: float, bias: float, valid: bool, train: bool):
        img1, img2 = _check(img1, "img1"), _check(img2, "img2")
        assert img1.dim() == 4 and img1.shape == img2.shape, (img1.shape, img2.shape)
        B, H, W, CH = img1.shape
        if B * H * W * CH == 0:
            raise ValueError("photometric loss of an empty image")
        dev = img1.device
        keep = train and ctx.needs_input_grad[0]
        # derivative maps for the backward; not needed for inference or for a pure L1 loss (coef_ssim == 0)
        maps = torch.empty((3, B, H, W, CH), dtype=torch.float32, device=dev) if keep and coef_ssim != 0 else None
        lib = _cabi.load_library()
        work = torch.empty(int(lib.gsb200_photometric_workspace_bytes()) // 8, dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        stats = torch.empty(2, dtype=torch.float32, device=dev)  # L1 mean, SSIM mean (for logging)
        with device_guard(dev):
            call("gsb200_photometric_loss_fwd", B, H, W, CH, ptr(img1), ptr(img2), float(coef_l1), float(coef_ssim), float(bias),
                 int(valid), ptr(None if maps is None else maps[0]), ptr(None if maps is None else maps[1]),
                 ptr(None if maps is None else maps[2]), ptr(work), ptr(out), ptr(stats), stream_ptr(dev))
        ctx.trained = keep
        if keep:
            ctx.save_for_backward(img1, img2, maps)
        ctx.coefs = (float(coef_l1), float(coef_ssim), bool(valid))
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, v_loss, _v_stats):
        if not ctx.trained:
            raise RuntimeError("photometric loss was evaluated with train=False; no gradient available")
        img1, img2, maps = ctx.saved_tensors
        coef_l1, coef_ssim, valid = ctx.coefs
        B, H, W, CH = img1.shape
        v_loss = v_loss.to(torch.float32).contiguous()
        v_img1 = torch.empty_like(img1)
        with device_guard(img1.device):
            m = (None, None, None) if maps is None else (maps[0], maps[1], maps[2])
            call("gsb200_photometric_loss_bwd", B, H, W, CH, ptr(img1), ptr(img2), coef_l1, coef_ssim, int(valid), ptr(m[0]),
                 ptr(m[1]), ptr(m[2]), ptr(v_loss), ptr(v_img1), stream_ptr(img1.device))
        return v_img1, None, None, None, None, None, None


def photometric_loss(colors: Tensor, pixels: Tensor, ssim_lambda: float = 0.2, padding: str = "valid",
                     return_stats: bool = False):
    """``F.l1_loss(colors, pixels) * (1 - ssim_lambda) + (1 - fused_ssim(colors, pixels, padding)) * ssim_lambda`` for
    [B,H,W,CH] images (examples/simple_trainer.py:948-953) in one forward and one backward kernel.
    With ``return_stats`` also returns a [2] tensor (L1 mean, SSIM mean) for logging.  ``ssim_lambda == 0`` is a
    pure L1 loss: the SSIM window is skipped altogether (the SSIM statistic is then reported as 0)."""
    assert padding in ("valid", "same"), padding
    loss, stats = _Photometric.apply(colors, pixels, 1.0 - ssim_lambda, -ssim_lambda, ssim_lambda, padding == "valid", True)
    return (loss, stats) if return_stats else loss


def _as_nhwc(img: Tensor) -> Tensor:
    """[B,CH,H,W] -> [B,H,W,CH] memory; free when `img` is a permuted view of channels-last data (the trainers pass
    ``colors.permute(0, 3, 1, 2)``), one copy otherwise."""
    assert img.dim() == 4, img.shape
    return img.permute(0, 2, 3, 1)


def fused_ssim(img1: Tensor, img2: Tensor, padding: str = "same", train: bool = True) -> Tensor:
    """Mean SSIM of [B,CH,H,W] images (11-tap Gaussian window, sigma 1.5), gradient to ``img1`` only -- the call
    signature of the ``fused_ssim`` package the reference trainers import."""
    assert padding in ("valid", "same"), padding
    value, _ = _Photometric.apply(_as_nhwc(img1), _as_nhwc(img2), 0.0, 1.0, 0.0, padding == "valid", train)
    return value


__all__ = ["fused_ssim", "photometric_loss"]
