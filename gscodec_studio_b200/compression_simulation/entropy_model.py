"""Factorized-prior entropy model on one fused CUDA kernel (libgsb200: gsb200_entropy_fwd / _bwd).

Class names, constructor arguments, parameter names (`_matrices.i`, `_bias.i`, `_factor.i` -- they
are written into checkpoints, examples/simple_trainer.py:1092-1095) and outputs follow the reference's
``Entropy_factorized_optimized_refactor`` (compression_simulation/entropy_model.py:195-308), including
its channel-scrambled weight tiling.  The reference's two older classes
contain live ``pdb.set_trace()`` calls (entropy_model.py:54,136,156) and are unusable; their names are
kept importable as aliases of the working implementation.

Activated-parameter layout handed to the kernel, per parameter set p (see csrc/quant.cu):
    layer 0: softplus(M0)[p] (3x1) | b0[p] (3) | tanh(f0)[p] (3)
    layer l: softplus(Ml)[p] (3x3 row-major) | bl[p] (3) | tanh(fl)[p] (3)      for 1 <= l < n_hidden
    last   : softplus(ML)[p] (1x3) | bL[p] (1)
The softplus / tanh and their chain rule run in torch on these 28- or 43-float tensors per channel.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as nnf
from torch import Tensor

from .._cabi import aligned, call, device_guard, ptr, stream_ptr


class _EntropyBits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, params: Tensor, q: float, n_hidden: int, bound: float, want_likelihood: bool):
        if not x.is_cuda:
            raise RuntimeError("gscodec_studio_b200: the entropy model needs CUDA tensors (no CPU implementation)")
        x = aligned(x.to(torch.float32))
        params = aligned(params.to(torch.float32))
        N, ch = x.shape
        bits = torch.empty_like(x)
        with device_guard(x.device):
            call("gsb200_entropy_fwd", N, ch, n_hidden, ptr(x), float(q), ptr(params), float(bound), ptr(bits),
                 stream_ptr(x.device))
        ctx.save_for_backward(x, params)
        ctx.cfg = (float(q), n_hidden, float(bound))
        return bits

    @staticmethod
    def backward(ctx, v_bits: Tensor):
        x, params = ctx.saved_tensors
        q, n_hidden, bound = ctx.cfg
        N, ch = x.shape
        v_bits = aligned(v_bits.contiguous())
        v_x = torch.empty_like(x)
        v_params = torch.zeros_like(params)
        with device_guard(x.device):
            call("gsb200_entropy_bwd", N, ch, n_hidden, ptr(x), q, ptr(params), bound, ptr(v_bits), ptr(v_x),
                 ptr(v_params), stream_ptr(x.device))
        return v_x, v_params, None, None, None, None


def _ptr_array(tensors):
    import ctypes

    arr = (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


class _EntropyParams(torch.autograd.Function):
    """softplus(matrices) | bias | tanh(factor) of every layer -> the kernels' [channel, PS] table."""

    @staticmethod
    def forward(ctx, n_hidden: int, channel: int, *raw):
        nl = n_hidden + 1
        raw = [t.detach().to(torch.float32).contiguous() for t in raw]
        mats, bias, fac = raw[:nl], raw[nl:2 * nl], raw[2 * nl:]
        ps = 9 + (n_hidden - 1) * 15 + 4
        table = torch.empty((channel, ps), dtype=torch.float32, device=mats[0].device)
        keep = (_ptr_array(mats), _ptr_array(bias), _ptr_array(fac))
        with device_guard(table.device):
            call("gsb200_entropy_params_fwd", channel, n_hidden, keep[0][1], keep[1][1], keep[2][1], ptr(table),
                 stream_ptr(table.device))
        ctx.save_for_backward(*mats, *fac)
        ctx.cfg = (n_hidden, channel)
        return table

    @staticmethod
    def backward(ctx, v_table):
        n_hidden, channel = ctx.cfg
        nl = n_hidden + 1
        saved = ctx.saved_tensors
        mats, fac = list(saved[:nl]), list(saved[nl:])
        v_table = v_table.contiguous()
        need = ctx.needs_input_grad[2:]
        v_m = [torch.empty_like(t) if need[i] else None for i, t in enumerate(mats)]
        v_b = [torch.empty((channel, t.shape[1], 1), dtype=torch.float32, device=t.device) if need[nl + i] else None
               for i, t in enumerate(mats)]
        v_f = [torch.empty_like(t) if need[2 * nl + i] else None for i, t in enumerate(fac)]
        keep = (_ptr_array(mats), _ptr_array(fac), _ptr_array(v_m), _ptr_array(v_b), _ptr_array(v_f))
        with device_guard(v_table.device):
            call("gsb200_entropy_params_bwd", channel, n_hidden, keep[0][1], keep[1][1], ptr(v_table), keep[2][1], keep[3][1],
                 keep[4][1], stream_ptr(v_table.device))
        return (None, None, *v_m, *v_b, *v_f)


class Entropy_factorized_optimized_refactor(nn.Module):
    def __init__(self, channel=32, init_scale=10, filters=(3, 3, 3), likelihood_bound=1e-6, tail_mass=1e-9,
                 optimize_integer_offset=True, Q=1):
        super().__init__()
        self.channel = channel
        self.filters = tuple(int(t) for t in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        self.optimize_integer_offset = bool(optimize_integer_offset)
        self.Q = Q
        self.likelihood_bound = float(likelihood_bound)
        if not 0 < self.tail_mass < 1:
            raise ValueError("`tail_mass` must be between 0 and 1")
        if self.filters not in ((3, 3), (3, 3, 3)):
            raise NotImplementedError("gscodec_studio_b200: the fused entropy kernel supports filters (3,3) and (3,3,3)")
        widths = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1.0 / (len(self.filters) + 1))
        self._matrices = nn.ParameterList()
        self._bias = nn.ParameterList()
        self._factor = nn.ParameterList()
        for i in range(len(self.filters) + 1):
            init = float(np.log(np.expm1(1.0 / scale / widths[i + 1])))
            self._matrices.append(nn.Parameter(torch.full((channel, widths[i + 1], widths[i]), init)))
            self._bias.append(nn.Parameter(torch.from_numpy(
                np.random.uniform(-0.5, 0.5, (channel, widths[i + 1], 1)).astype(np.float32))))
            if i < len(self.filters):
                self._factor.append(nn.Parameter(torch.zeros(channel, widths[i + 1], 1)))
        # checkpoint compatibility: the reference leaves the LAST layer's parameters registered on the module as well
        # (entropy_model.py:216-232 assigns self.matrix / self.bias / self.factor in its loop), so its state_dict carries
        # `matrix`, `bias`, `factor` (aliases of _matrices[-1], _bias[-1], _factor[-1]) and `likelihood_lower_bound.bound`;
        # the trainer saves and strictly re-loads that dict (examples/simple_trainer.py:1093-1095, 1474)
        self.matrix = self._matrices[-1]
        self.bias = self._bias[-1]
        self.factor = self._factor[-1]
        self.register_buffer("filters_len", torch.tensor(len(self.filters)))
        self.register_buffer("factor_len", torch.tensor(len(self._factor)))
        self.likelihood_lower_bound = LowerBound(likelihood_bound)

    def _activated_params(self) -> Tensor:
        """[channel, PS] tensor in the kernel's layout; differentiable w.r.t. the raw parameters.  One launch each way
        (gsb200_entropy_params_fwd / _bwd) instead of ~35 tiny torch kernels per model and step."""
        nl = len(self.filters) + 1
        raw = [self._matrices[i] for i in range(nl)] + [self._bias[i] for i in range(nl)] + \
              [self._factor[i] for i in range(nl - 1)]
        if not raw[0].is_cuda:
            raise RuntimeError("gscodec_studio_b200: the entropy model needs CUDA tensors (no CPU implementation)")
        return _EntropyParams.apply(nl - 1, self.channel, *raw)

    def _q_value(self, Q) -> float:
        if Q is None:
            Q = self.Q
        if isinstance(Q, Tensor):
            if Q.numel() != 1:
                raise NotImplementedError("gscodec_studio_b200: per-channel Q tensors are not supported by the fused kernel")
            Q = float(Q.item())
        return float(Q)

    def forward(self, x, Q=None, **kwargs):
        """x [N, channel] (fake-quantised values), Q the quantisation step -> estimated bits [N, channel]."""
        return _EntropyBits.apply(x, self._activated_params(), self._q_value(Q), len(self.filters),
                                  self.likelihood_bound, False)

    def get_likelihood(self, x, Q=None, **kwargs):
        """Likelihoods [N, channel] (entropy_model.py:256-308): 2^-bits of the same kernel."""
        with torch.no_grad():
            bits = _EntropyBits.apply(x, self._activated_params(), self._q_value(Q), len(self.filters),
                                      self.likelihood_bound, True)
        return torch.exp2(-bits)


# the reference's older classes are dead code (pdb.set_trace in forward); keep the names importable
Entropy_factorized = Entropy_factorized_optimized_refactor
Entropy_factorized_optimized = Entropy_factorized_optimized_refactor


class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, bound):
        ctx.save_for_backward(inputs)
        ctx.bound = bound
        return torch.clamp_min(inputs, bound)

    @staticmethod
    def backward(ctx, grad_output):
        (inputs,) = ctx.saved_tensors
        passes = (inputs >= ctx.bound) | (grad_output < 0)  # the bound still lets the value be pushed up
        return passes.type(grad_output.dtype) * grad_output, None


class LowerBound(nn.Module):
    """max(x, bound) whose gradient still flows when it moves x towards the feasible side (reference
    entropy_model.py:347-390); used by the hash-grid model only -- the factorized model has it fused in its kernel."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))  # a buffer, as in the reference's state_dict
        self._bound = float(bound)

    def forward(self, x):
        return _LowerBoundFn.apply(x, self._bound)


class Entropy_gaussian(nn.Module):
    """Hash-grid conditioned Gaussian entropy model (reference entropy_model.py:313-345): a hash-grid + MLP regressor
    predicts a per-channel (mean, scale) from the Gaussian's position, the rate is -log2 of the Normal mass of the
    quantisation bin.  The grid encode / decode is csrc/hashgrid.cu, validated on B200 against a CPU restatement
    and, live, against the reference's unmodified `_gridencoder` extension (tests/test_gpu_widen.py)."""

    def __init__(self, channel=3, Q=1, likelihood_bound=1e-6):
        super().__init__()
        from .gaussian_distribution_model import hash_based_estimator

        self.Q = Q
        self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.mean, self.scale = 0, 1
        self.param_regressor = hash_based_estimator(channel)

    def forward(self, x, Q=None, pos=None):
        assert pos is not None
        assert x.size()[0] == pos.size()[0]
        self.mean, self.scale = self.param_regressor(pos)
        self.scale = torch.clamp(self.scale, min=1e-9)
        normal = torch.distributions.normal.Normal(self.mean, self.scale)
        if Q is None:
            Q = self.Q
        likelihood = torch.abs(normal.cdf(x + 0.5 * Q) - normal.cdf(x - 0.5 * Q))
        return -torch.log2(self.likelihood_lower_bound(likelihood))

    def get_means_and_scales(self, pos):
        means, scales = self.param_regressor(pos)
        return means, torch.clamp(scales, min=1e-9)
