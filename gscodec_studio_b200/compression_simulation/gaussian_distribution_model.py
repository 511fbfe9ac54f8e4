"""Hash-grid parameter regressor of the alternative in-loop entropy model.

Mirrors the class names, constructor arguments and ``state_dict`` keys of the reference's
``gsplat/compression_simulation/gaussian_distribution_model.py`` (``GridEncoder`` :184-274, ``mix_3D2D_encoding``
:276-346, ``hash_based_estimator`` :348-382) so that checkpoints are interchangeable.  The grid encode / decode runs in
the CUDA kernels of ``csrc/hashgrid.cu`` (STE-binary transform fused); the two small ``nn.Linear`` layers stay torch.
The kernels are tested on B200 against a CPU restatement and, live, against the reference's CUDA-only ``_gridencoder``
extension built unmodified (tests/test_gpu_widen.py; DESIGN.md section 2 has the pinning numbers).
"""
from __future__ import annotations

import ctypes
import math
from typing import Sequence

import torch
import torch.nn as nn
from torch import Tensor

from .._cabi import call, device_guard, ptr, stream_ptr


class _GridEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs: Tensor, params: Tensor, offsets: Sequence[int], resolutions: Sequence[int], ste_binary: bool):
        if not inputs.is_cuda or not params.is_cuda:
            raise RuntimeError("gscodec_studio_b200: the hash-grid encoder needs CUDA tensors (no CPU implementation)")
        inputs, params = inputs.float().contiguous(), params.float().contiguous()
        N, D = inputs.shape
        F, L = params.shape[1], len(resolutions)
        offs = (ctypes.c_int32 * (L + 1))(*[int(o) for o in offsets])
        ress = (ctypes.c_int32 * L)(*[int(r) for r in resolutions])
        out = torch.empty((N, L * F), dtype=torch.float32, device=inputs.device)
        with device_guard(inputs.device):
            call("gsb200_grid_encode_fwd", N, D, F, L, ctypes.cast(offs, ctypes.c_void_p), ctypes.cast(ress, ctypes.c_void_p),
                 int(ste_binary), ptr(inputs), ptr(params), ptr(out), stream_ptr(inputs.device))
        ctx.save_for_backward(inputs, params)
        ctx.meta = (tuple(int(o) for o in offsets), tuple(int(r) for r in resolutions), bool(ste_binary))
        return out

    @staticmethod
    def backward(ctx, v_out: Tensor):
        inputs, params = ctx.saved_tensors
        offsets, resolutions, ste_binary = ctx.meta
        N, D = inputs.shape
        F, L = params.shape[1], len(resolutions)
        offs = (ctypes.c_int32 * (L + 1))(*offsets)
        ress = (ctypes.c_int32 * L)(*resolutions)
        v_out = v_out.float().contiguous()
        v_params = torch.empty_like(params)
        v_inputs = torch.empty_like(inputs) if ctx.needs_input_grad[0] else None
        with device_guard(inputs.device):
            call("gsb200_grid_encode_bwd", N, D, F, L, ctypes.cast(offs, ctypes.c_void_p), ctypes.cast(ress, ctypes.c_void_p),
                 int(ste_binary), ptr(inputs), ptr(params), ptr(v_out), ptr(v_params), ptr(v_inputs), stream_ptr(inputs.device))
        return v_inputs, v_params, None, None, None


class GridEncoder(nn.Module):
    def __init__(self, num_dim=3, n_features=2, resolutions_list=(16, 23, 32, 46, 64, 92, 128, 184, 256, 368, 512, 736),
                 log2_hashmap_size=19, ste_binary=True, ste_multistep=False, add_noise=False, Q=1):
        super().__init__()
        if ste_multistep or add_noise:
            raise NotImplementedError("GridEncoder: only the ste_binary / plain embedding modes the trainers use are built")
        if n_features != 2 or num_dim not in (2, 3):
            raise NotImplementedError("GridEncoder: 2 features per level and 2 or 3 input dimensions")
        res = torch.tensor(resolutions_list).to(torch.int)
        self.num_dim, self.n_levels, self.n_features = num_dim, res.numel(), n_features
        self.log2_hashmap_size, self.ste_binary, self.Q = log2_hashmap_size, ste_binary, Q
        self.output_dim = self.n_output_dims = self.n_levels * n_features
        offsets, offset = [], 0
        for r in res.tolist():
            n = min(2 ** log2_hashmap_size, r ** num_dim)
            n = int(math.ceil(n / 8) * 8)
            offsets.append(offset)
            offset += n
        offsets.append(offset)
        self.register_buffer("offsets_list", torch.tensor(offsets, dtype=torch.int32))
        self.register_buffer("resolutions_list", res)
        self._offsets, self._resolutions = offsets, res.tolist()  # host copies: the kernels take them by value
        self.params = nn.Parameter(torch.empty(offset, n_features))
        self.reset_parameters()

    def reset_parameters(self):
        self.params.data.uniform_(-1e-4, 1e-4)

    def forward(self, inputs: Tensor) -> Tensor:
        prefix = list(inputs.shape[:-1])
        out = _GridEncode.apply(inputs.reshape(-1, self.num_dim), self.params, self._offsets, self._resolutions, self.ste_binary)
        return out.view(prefix + [self.output_dim])


class mix_3D2D_encoding(nn.Module):
    def __init__(self, n_features, resolutions_list, log2_hashmap_size, resolutions_list_2D, log2_hashmap_size_2D, ste_binary,
                 ste_multistep, add_noise, Q):
        super().__init__()
        kw = dict(n_features=n_features, ste_binary=ste_binary, ste_multistep=ste_multistep, add_noise=add_noise, Q=Q)
        self.encoding_xyz = GridEncoder(num_dim=3, resolutions_list=resolutions_list, log2_hashmap_size=log2_hashmap_size, **kw)
        self.encoding_xy = GridEncoder(num_dim=2, resolutions_list=resolutions_list_2D, log2_hashmap_size=log2_hashmap_size_2D, **kw)
        self.encoding_xz = GridEncoder(num_dim=2, resolutions_list=resolutions_list_2D, log2_hashmap_size=log2_hashmap_size_2D, **kw)
        self.encoding_yz = GridEncoder(num_dim=2, resolutions_list=resolutions_list_2D, log2_hashmap_size=log2_hashmap_size_2D, **kw)
        self.output_dim = (self.encoding_xyz.output_dim + self.encoding_xy.output_dim + self.encoding_xz.output_dim +
                           self.encoding_yz.output_dim)

    def forward(self, x: Tensor) -> Tensor:
        xx, yy, zz = torch.chunk(x, 3, dim=-1)
        return torch.cat([self.encoding_xyz(x), self.encoding_xy(torch.cat([xx, yy], dim=-1)),
                          self.encoding_xz(torch.cat([xx, zz], dim=-1)), self.encoding_yz(torch.cat([yy, zz], dim=-1))], dim=-1)


class hash_based_estimator(nn.Module):
    def __init__(self, channel=4):
        super().__init__()
        self.channel = channel
        self.hash_grid = mix_3D2D_encoding(
            n_features=2, resolutions_list=(18, 24, 33, 44, 59, 80, 108, 148, 201, 275, 376, 514), log2_hashmap_size=13,
            resolutions_list_2D=(130, 258, 514, 1026), log2_hashmap_size_2D=15, ste_binary=True, ste_multistep=False,
            add_noise=False, Q=1)
        self.mlp_regressor = nn.Sequential(nn.Linear(self.hash_grid.output_dim, channel * 5), nn.ReLU(True),
                                           nn.Linear(channel * 5, channel * 2))
        self.bound_min, self.bound_max = -100.0, 100.0

    def forward(self, x: Tensor):
        assert len(x.shape) == 2 and x.shape[1] == 3
        x = (x - self.bound_min) / (self.bound_max - self.bound_min)
        out = self.mlp_regressor(self.hash_grid(x))
        return out[:, 0:self.channel], out[:, self.channel:2 * self.channel]
