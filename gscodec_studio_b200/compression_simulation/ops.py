"""Fake-quantisation ops, API of the reference's ``gsplat/compression_simulation/ops.py`` (:14-75) on one
fused CUDA kernel each way (libgsb200: gsb200_fake_quantize_fwd / _bwd).

Additive keyword arguments (ignored by reference callers):
  noise       explicit U(-.5,.5) tensor for the "noise" mode (parity tests inject the reference's draw);
              default: drawn in-kernel with Philox from torch's CUDA generator state (seed, offset)
  activation  None | "exp" | "sigmoid": fuse the activation the trainer applies next
              (examples/simple_trainer.py:758-776) into the same pass.

`QuantOnLoad` (additive) is the tagged, lazy form of a noise-mode result: `rasterization()` recognises it and applies the
quantisation inside the projection / colour kernels, where the raw parameter is loaded; every other consumer sees the
materialised tensor (same values: the uniforms come from the same Philox counters or the same injected noise).
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
    def forward(ctx, x: Tensor, lo: float, hi: float, bitwidth: int, mode: int, act: int, noise: Optional[Tensor],
                philox: Optional[tuple] = None):
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
            seed, offset = philox if philox is not None else _philox_args(x.device, x.numel())
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
            return v_y, None, None, None, None, None, None, None  # pure straight-through (ops.py:73-75)
        v_y = aligned(v_y.contiguous())
        v_x = torch.empty_like(x)
        with device_guard(x.device):
            call("gsb200_fake_quantize_bwd", x.numel(), ptr(x), ptr(y), lo, hi, mode, act, ptr(v_y), ptr(v_x),
                 stream_ptr(x.device))
        return v_x.view_as(v_y), None, None, None, None, None, None, None


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


class QuantOnLoad:
    """Noise-mode ``fake_quantize_ste(raw, lo, hi, bitwidth)["output_value"]`` that has not been computed yet.

    The fusion hook of SURVEY.md section 8b: ``simulate_compression(..., fuse_into_render=True)`` returns these for the
    attributes it only fake-quantises; ``torch.exp`` / ``torch.sigmoid`` of one is again lazy (the activation is recorded);
    ``rasterization()`` hands raw parameter + descriptor to kernels that quantise on load (gsb200_projection_quant_*,
    gsb200_sh_view_split_quant_*).  Anything else -- any other torch function, attribute, arithmetic, indexing -- works on
    the materialised tensor, computed once by the stand-alone kernel with the SAME uniforms, so both views agree bit for bit.
    """

    __slots__ = ("raw", "lo", "hi", "bitwidth", "noise", "seed", "offset", "act", "_value", "_family")

    def __init__(self, raw: Tensor, lower_bd: float, upper_bd: float, bitwidth: int = 8, noise: Optional[Tensor] = None,
                 activation: Optional[str] = None, philox: Optional[tuple] = None):
        if raw.dtype != torch.float32:
            raise RuntimeError(f"gscodec_studio_b200: expected float32 tensor, got {raw.dtype}")
        self.raw = raw  # (a CPU tensor is refused where a kernel would touch it: there is no CPU implementation)
        self.lo, self.hi, self.bitwidth = float(lower_bd), float(upper_bd), int(bitwidth)
        self.act = None if activation in (None, "none") else activation
        if self.act not in (None, "exp", "sigmoid"):
            raise ValueError(f"unknown activation {activation!r}")
        self.noise, self.seed, self.offset = None, 0, 0
        if noise is not None:
            assert noise.numel() == raw.numel(), "noise must have as many elements as the input"
            self.noise = aligned(noise.to(torch.float32))
        else:
            self.seed, self.offset = philox if philox is not None else _philox_args(raw.device, raw.numel())
        self._value = None
        self._family = {}  # activation -> lazy sibling (shares the draw)

    # ---- what rasterization() needs
    @property
    def q_step(self) -> float:
        return (self.hi - self.lo) / (2 ** self.bitwidth - 1)

    def descriptor(self):
        """ctypes struct gsb200_quant_load for one kernel call (the caller keeps `self` alive for the noise pointer)."""
        from .._cabi import QuantLoadDesc

        d = QuantLoadDesc()
        d.noise = self.noise.data_ptr() if self.noise is not None else None
        d.seed, d.offset, d.lo, d.hi, d.bitwidth, d.act = self.seed, self.offset, self.lo, self.hi, self.bitwidth, _ACT[self.act]
        return d

    def raw_tensor(self) -> Tensor:
        """The raw parameter as the kernels read it (contiguous, 16-byte aligned; autograd-connected to `raw`)."""
        return aligned(self.raw)

    def materialize(self) -> Tensor:
        if self._value is None:
            philox = None if self.noise is not None else (self.seed, self.offset)
            self._value = _FakeQuantize.apply(self.raw, self.lo, self.hi, self.bitwidth, 0, _ACT[self.act], self.noise, philox)
        return self._value

    def _activated(self, act: str):
        if self.act is not None:  # an activation of an activation: no longer an attribute load
            return getattr(torch, act)(self.materialize())
        sib = self._family.get(act)
        if sib is None:
            sib = QuantOnLoad.__new__(QuantOnLoad)
            sib.raw, sib.lo, sib.hi, sib.bitwidth = self.raw, self.lo, self.hi, self.bitwidth
            sib.noise, sib.seed, sib.offset, sib.act, sib._value, sib._family = self.noise, self.seed, self.offset, act, None, {}
            self._family[act] = sib
        return sib

    def exp(self):
        return self._activated("exp")

    def sigmoid(self):
        return self._activated("sigmoid")

    # ---- tensor-likeness without computing anything
    shape = property(lambda self: self.raw.shape)
    device = property(lambda self: self.raw.device)
    dtype = property(lambda self: self.raw.dtype)
    requires_grad = property(lambda self: self.raw.requires_grad)
    is_cuda = property(lambda self: self.raw.is_cuda)
    ndim = property(lambda self: self.raw.ndim)

    def dim(self):
        return self.raw.dim()

    def size(self, *a):
        return self.raw.size(*a)

    def numel(self):
        return self.raw.numel()

    def __len__(self):
        return len(self.raw)

    # ---- everything else: the materialised tensor
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (torch.exp, torch.sigmoid) and len(args) == 1 and not kwargs and isinstance(args[0], cls):
            return args[0]._activated(func.__name__)

        def solid(v):
            if isinstance(v, cls):
                return v.materialize()
            if isinstance(v, (list, tuple)):
                return type(v)(solid(e) for e in v)
            return v

        return func(*solid(args), **{k: solid(v) for k, v in kwargs.items()})

    def __getattr__(self, name):  # only reached for names not defined above
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __neg__(self):
        return -self.materialize()

    def __repr__(self):
        return (f"QuantOnLoad(shape={tuple(self.raw.shape)}, bounds=({self.lo}, {self.hi}), bitwidth={self.bitwidth}, "
                f"activation={self.act}, materialised={self._value is not None})")


def _binary(name):
    def op(self, other):
        other = other.materialize() if isinstance(other, QuantOnLoad) else other
        return getattr(self.materialize(), name)(other)

    op.__name__ = name
    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__", "__pow__",
           "__matmul__", "__lt__", "__le__", "__gt__", "__ge__"):
    setattr(QuantOnLoad, _n, _binary(_n))


def solid(value):
    """The tensor behind `value` (a Tensor, or a QuantOnLoad that gets materialised); None stays None."""
    return value.materialize() if isinstance(value, QuantOnLoad) else value


class STE(torch.autograd.Function):
    """Round-mode straight-through estimator with the reference's signature (ops.py:57-75)."""

    @staticmethod
    def forward(ctx, input, bitdepth=8, min=-1, max=1):
        return _FakeQuantize.apply(input, min, max, bitdepth, 1, 0, None, None)

    @staticmethod
    def backward(ctx, grad_output):  # pragma: no cover - apply() of the inner Function records the graph
        return grad_output, None, None, None


def fake_quantize_ste(input: Tensor, lower_bd: float, upper_bd: float, bitwidth: int = 8, q_type: str = "noise",
                      noise: Optional[Tensor] = None, activation: Optional[str] = None) -> Dict:
    """-> {"output_value": Tensor, "q_step": float}; q_type "noise" (default) or "round"."""
    q_step = (upper_bd - lower_bd) / (2**bitwidth - 1)
    if q_type not in _MODE:
        raise NotImplementedError(f"q_type {q_type!r} is not supported")
    out = _FakeQuantize.apply(input, lower_bd, upper_bd, bitwidth, _MODE[q_type], _ACT[activation], noise, None)
    return {"output_value": out, "q_step": q_step}


def fake_quantize(input: Tensor, lower_bd: float, upper_bd: float, bitwidth: int = 8, q_type: str = "round") -> Dict:
    """Reference ops.py:14-37: clamp + round with the (fq - x).detach() + x straight-through form, which
    does NOT mutate the input.  Built from the round kernel on a copy."""
    q_step = (upper_bd - lower_bd) / (2**bitwidth - 1)
    if q_type != "round":
        raise NotImplementedError
    with torch.no_grad():
        fq = _FakeQuantize.apply(input.detach().clone(), lower_bd, upper_bd, bitwidth, 1, 0, None, None)
    return {"output_value": (fq - input).detach() + input, "q_step": q_step}
