"""Optimizers next to the hot path.

``SelectiveAdam`` keeps the API of the reference's ``gsplat/optimizers/selective_adam.py:7-94`` (an Adam
whose ``step(visibility)`` updates only the rows flagged visible, one fused kernel per parameter).

``GroupedAdam`` is additive: ``torch.optim.Adam`` semantics (the optimizer the trainers build per attribute,
examples/simple_trainer.py:396-408) with ONE kernel launch for all parameter groups -- every element of
(param, grad, exp_avg, exp_avg_sq) is read once and (param, exp_avg, exp_avg_sq) written once.  Its state
dict has torch.optim.Adam's keys (``step``, ``exp_avg``, ``exp_avg_sq``), so checkpoints are interchangeable.
"""
from __future__ import annotations

import ctypes
import math

import torch

from .. import _cabi
from .._cabi import call, device_guard, ptr, stream_ptr
from ..cuda._wrapper import selective_adam_update


def _moments(optimizer: torch.optim.Optimizer, param: torch.Tensor):
    """The (lazily created) Adam state of `param`, laid out like torch.optim.Adam's: step, exp_avg, exp_avg_sq."""
    state = optimizer.state[param]
    if not state:
        state["step"] = torch.tensor(0.0, dtype=torch.float32)
        for key in ("exp_avg", "exp_avg_sq"):
            state[key] = torch.zeros_like(param, memory_format=torch.preserve_format)
    return state


class SelectiveAdam(torch.optim.Adam):
    """Adam whose ``step(visibility)`` touches only the rows flagged in ``visibility`` (bool [N]) -- parameter AND
    moments of the other rows stay as they are -- with one fused kernel per parameter tensor (the second of the two
    optimizers of the Taming-3DGS paper, as in the reference).  One tensor per parameter group; no bias correction.

    >>> optimizer = SelectiveAdam([param], eps=1e-8, betas=(0.9, 0.999))
    >>> loss.backward()
    >>> optimizer.step(visibility=visibility_mask)
    """

    def __init__(self, params, eps, betas):
        super().__init__(params=params, eps=eps, betas=betas)

    @torch.no_grad()
    def step(self, visibility):
        n_rows = visibility.numel()
        for group in self.param_groups:
            assert len(group["params"]) == 1, "more than one tensor in group"
            (param,) = group["params"]
            if param.grad is None:
                continue
            state = _moments(self, param)
            beta1, beta2 = group["betas"]
            selective_adam_update(param, param.grad, state["exp_avg"], state["exp_avg_sq"], visibility, group["lr"], beta1,
                                  beta2, group["eps"], n_rows, param.numel() // n_rows)


class GroupedAdam(torch.optim.Adam):
    """torch.optim.Adam (no weight decay, no amsgrad, not maximize) for all parameter groups in one launch."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params=params, lr=lr, betas=betas, eps=eps)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        rows = []
        keep = []  # tensors that must stay alive until the launch has been enqueued
        device = None
        for group in self.param_groups:
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise NotImplementedError("GroupedAdam: weight_decay / amsgrad / maximize are not supported")
            lr, eps = float(group["lr"]), float(group["eps"])
            beta1, beta2 = group["betas"]
            for param in group["params"]:
                if param.grad is None:
                    continue
                if not param.is_cuda or param.dtype != torch.float32:
                    raise RuntimeError("GroupedAdam: parameters must be float32 CUDA tensors (no CPU implementation)")
                if param.grad.is_sparse:
                    raise RuntimeError("GroupedAdam does not support sparse gradients")
                if not param.is_contiguous():
                    raise RuntimeError("GroupedAdam: parameters must be contiguous")
                device = param.device if device is None else device
                if param.device != device:
                    raise RuntimeError("GroupedAdam: all parameters must live on one device")
                state = _moments(self, param)
                state["step"] += 1
                t = float(state["step"])
                grad = param.grad.contiguous()
                keep.append(grad)
                rows.append(_cabi.AdamTensor(
                    ptr(param), ptr(grad), ptr(state["exp_avg"]), ptr(state["exp_avg_sq"]), param.numel(),
                    lr / (1.0 - beta1 ** t), beta1, beta2, eps, 1.0 / math.sqrt(1.0 - beta2 ** t), 1.0 - beta1, 1.0 - beta2))
        with device_guard(device) if device is not None else _null():
            for i in range(0, len(rows), _cabi.ADAM_MAX_TENSORS):
                part = rows[i:i + _cabi.ADAM_MAX_TENSORS]
                table = (_cabi.AdamTensor * len(part))(*part)
                call("gsb200_adam_multi", len(part), ctypes.cast(table, ctypes.c_void_p), stream_ptr(device))
        return loss


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
