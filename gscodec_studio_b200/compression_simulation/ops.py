"""Fake-quantisation ops, API of the reference's ``gsplat/compression_simulation/ops.py`` (:14-75) on one
fused CUDA kernel each way (libgsb200: gsb200_fake_quantize_fwd / _bwd).

Additive keyword arguments (ignored by reference callers):
  noise       explicit U(-.5,.5) tensor for the "noise" mode (parity tests inject the reference's draw);
              default: drawn in-kernel with Philox from torch's CUDA generator state (seed, offset)
  activation  None | "exp" | "sigmoid": fuse the activation the trainer applies next
              (examples/simple_trainer.py:758-776) into the same pass.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from .._cabi import aligned, call, device_guard, ptr, stream_ptr

_ACT = {None: 0, "none": 0, "exp": 1, "sigmoid": 2}
_MODE = {"noise": 0, "round": 1}


def log_transform(x):
    return torch.sign(x) * torch.log1p(torch.abs(x))


def inverse_log_transform(y):
    return torch.sign(y) * (torch.expm1(torch.abs(y)))


def _philox_args(device, n: int) -> tuple:
    """(seed, offset) taken from -- and advancing -- torch's CUDA generator, so runs are reproducible
    under torch.manual_seed like the reference's uniform_() draw (ops.py:46)."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed = gen.initial_seed() & (2**64 - 1)
    offset = gen.get_offset()
    gen.set_offset(offset + (n + 3) // 4 * 4)  # one Philox counter per 4 elements; keep the offset 4-aligned
    return seed, offset // 4


class _FakeQuantize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, lo: float, hi: float, bitwidth: int, mode: int, act: int, noise: Optional[Tensor]):
        if not x.is_cuda:
            raise RuntimeError("gscodec_studio_b200: fake_quantize needs a CUDA tensor (no CPU implementation)")
        if x.dtype != torch.float32:
            raise RuntimeError(f"gscodec_studio_b200: expected float32 tensor, got {x.dtype}")
        if mode == 1:
            # "round" clamps the parameter in place (reference ops.py:63); it must be written through `x`
            if not x.is_contiguous():
                raise RuntimeError("gscodec_studio_b200: round-mode fake quantisation needs a contiguous tensor")
            xin = x
        else:
            xin = aligned(x)
        y = torch.empty_like(xin)
        seed, offset = (0, 0)
        if mode == 0 and noise is None:
            seed, offset = _philox_args(x.device, x.numel())
        if noise is not None:
            noise = aligned(noise.to(torch.float32))
            assert noise.numel() == x.numel(), "noise must have as many elements as the input"
        with device_guard(x.device):
            call("gsb200_fake_quantize_fwd", xin.numel(), ptr(xin), float(lo), float(hi), int(bitwidth), mode, act,
                 ptr(noise), seed, offset, ptr(y), stream_ptr(x.device))
        # round mode wrote the clamped values through `x` (like the reference's input.clamp_ inside
        # STE.forward, which also does not tell autograd about it)
        ctx.save_for_backward(xin if mode == 0 else xin.detach(), y)
        ctx.cfg = (float(lo), float(hi), mode, act)
        return y.view_as(x)

    @staticmethod
    def backward(ctx, v_y: Tensor):
        x, y = ctx.saved_tensors
        lo, hi, mode, act = ctx.cfg
        if mode == 1 and act == 0:
            return v_y, None, None, None, None, None, None  # pure straight-through (ops.py:73-75)
        v_y = aligned(v_y.contiguous())
        v_x = torch.empty_like(x)
        with device_guard(x.device):
            call("gsb200_fake_quantize_bwd", x.numel(), ptr(x), ptr(y), lo, hi, mode, act, ptr(v_y), ptr(v_x),
                 stream_ptr(x.device))
        return v_x.view_as(v_y), None, None, None, None, None, None


class _FakeQuantizeMulti(torch.autograd.Function):
    """Several tensors quantised by ONE launch each way (gsb200_fake_quantize_multi_fwd / _bwd): the policy quantises
    4-7 attributes per training step and at 1 M Gaussians a launch per attribute costs more host time than device time.
    apply(cfgs, *tensors) with cfgs = [(lo, hi, bitwidth, mode, act, noise-or-None), ...] -> tuple of outputs."""

    @staticmethod
    def forward(ctx, cfgs, *xs):
        from .._cabi import FqTensor

        dev = xs[0].device
        table = (FqTensor * len(xs))()
        ins, outs, keep = [], [], []
        for i, (x, (lo, hi, bits, mode, act, noise)) in enumerate(zip(xs, cfgs)):
            if not x.is_cuda:
                raise RuntimeError("gscodec_studio_b200: fake_quantize needs a CUDA tensor (no CPU implementation)")
            if x.dtype != torch.float32:
                raise RuntimeError(f"gscodec_studio_b200: expected float32 tensor, got {x.dtype}")
            if mode == 1:
                if not x.is_contiguous():
                    raise RuntimeError("gscodec_studio_b200: round-mode fake quantisation needs a contiguous tensor")
                xin = x
            else:
                xin = aligned(x)
            y = torch.empty_like(xin)
            seed, offset = (0, 0)
            if mode == 0 and noise is None:
                seed, offset = _philox_args(dev, x.numel())
            if noise is not None:
                noise = aligned(noise.to(torch.float32))
                assert noise.numel() == x.numel(), "noise must have as many elements as the input"
                keep.append(noise)
            t = table[i]
            t.n, t.x, t.y, t.noise = xin.numel(), ptr(xin), ptr(y), ptr(noise)
            t.seed, t.offset, t.lo, t.hi, t.bitwidth, t.mode, t.act = seed, offset, float(lo), float(hi), int(bits), mode, act
            ins.append(xin if mode == 0 else xin.detach())
            outs.append(y)
        import ctypes as _ct

        with device_guard(dev):
            call("gsb200_fake_quantize_multi_fwd", len(xs), _ct.cast(table, _ct.c_void_p), stream_ptr(dev))
        ctx.save_for_backward(*ins, *outs)
        ctx.cfgs = [(float(lo), float(hi), int(bits), mode, act) for (lo, hi, bits, mode, act, _n) in cfgs]
        return tuple(y.view_as(x) for y, x in zip(outs, xs))

    @staticmethod
    def backward(ctx, *v_ys):
        from .._cabi import FqTensor
        import ctypes as _ct

        k = len(ctx.cfgs)
        ins, outs = ctx.saved_tensors[:k], ctx.saved_tensors[k:]
        dev = ins[0].device
        table = (FqTensor * k)()
        grads, keep, m = [], [], 0
        for i, (x, y, v_y, (lo, hi, bits, mode, act)) in enumerate(zip(ins, outs, v_ys, ctx.cfgs)):
            if v_y is None or not ctx.needs_input_grad[1 + i]:
                grads.append(None)
                continue
            if mode == 1 and act == 0:
                grads.append(v_y)  # pure straight-through (ops.py:73-75)
                continue
            v_y = aligned(v_y.contiguous())
            v_x = torch.empty_like(x)
            t = table[m]
            t.n, t.x, t.y, t.v_y, t.v_x = x.numel(), ptr(x), ptr(y), ptr(v_y), ptr(v_x)
            t.lo, t.hi, t.bitwidth, t.mode, t.act = lo, hi, bits, mode, act
            keep.append(v_y)
            grads.append(v_x.view_as(v_y))
            m += 1
        if m:
            with device_guard(dev):
                call("gsb200_fake_quantize_multi_bwd", m, _ct.cast(table, _ct.c_void_p), stream_ptr(dev))
        return (None, *grads)


def fake_quantize_ste_multi(items) -> list:
    """Additive: fake_quantize_ste for several tensors in one launch per direction.  items: sequence of dicts with the
    keyword arguments of fake_quantize_ste (input, lower_bd, upper_bd, bitwidth, q_type, noise, activation); returns the
    list of {"output_value", "q_step"} dicts in the same order."""
    from .._cabi import FQ_MAX_TENSORS

    items = list(items)
    out = []
    for lo_i in range(0, len(items), FQ_MAX_TENSORS):
        chunk = items[lo_i:lo_i + FQ_MAX_TENSORS]
        cfgs, xs = [], []
        for it in chunk:
            q_type = it.get("q_type", "noise")
            if q_type not in _MODE:
                raise NotImplementedError(f"q_type {q_type!r} is not supported")
            cfgs.append((it["lower_bd"], it["upper_bd"], it.get("bitwidth", 8), _MODE[q_type], _ACT[it.get("activation")],
                         it.get("noise")))
            xs.append(it["input"])
        ys = _FakeQuantizeMulti.apply(cfgs, *xs)
        for it, y in zip(chunk, ys):
            out.append({"output_value": y, "q_step": (it["upper_bd"] - it["lower_bd"]) / (2 ** it.get("bitwidth", 8) - 1)})
    return out


class STE(torch.autograd.Function):
    """Round-mode straight-through estimator with the reference's signature (ops.py:57-75)."""

    @staticmethod
    def forward(ctx, input, bitdepth=8, min=-1, max=1):
        return _FakeQuantize.apply(input, min, max, bitdepth, 1, 0, None)

    @staticmethod
    def backward(ctx, grad_output):  # pragma: no cover - apply() of the inner Function records the graph
        return grad_output, None, None, None


def fake_quantize_ste(input: Tensor, lower_bd: float, upper_bd: float, bitwidth: int = 8, q_type: str = "noise",
                      noise: Optional[Tensor] = None, activation: Optional[str] = None) -> Dict:
    """-> {"output_value": Tensor, "q_step": float}; q_type "noise" (default) or "round"."""
    q_step = (upper_bd - lower_bd) / (2**bitwidth - 1)
    if q_type not in _MODE:
        raise NotImplementedError(f"q_type {q_type!r} is not supported")
    out = _FakeQuantize.apply(input, lower_bd, upper_bd, bitwidth, _MODE[q_type], _ACT[activation], noise)
    return {"output_value": out, "q_step": q_step}


def fake_quantize(input: Tensor, lower_bd: float, upper_bd: float, bitwidth: int = 8, q_type: str = "round") -> Dict:
    """Reference ops.py:14-37: clamp + round with the (fq - x).detach() + x straight-through form, which
    does NOT mutate the input.  Built from the round kernel on a copy."""
    q_step = (upper_bd - lower_bd) / (2**bitwidth - 1)
    if q_type != "round":
        raise NotImplementedError
    with torch.no_grad():
        fq = _FakeQuantize.apply(input.detach().clone(), lower_bd, upper_bd, bitwidth, 1, 0, None)
    return {"output_value": (fq - input).detach() + input, "q_step": q_step}
