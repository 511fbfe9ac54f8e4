"""MCMC relocation op, API of the reference's ``gsplat/relocation.py:10-49``."""
from typing import Tuple

import torch
from torch import Tensor

from ._cabi import call, ptr, stream_ptr


def compute_relocation(
    opacities: Tensor,  # [N]
    scales: Tensor,  # [N, 3]
    ratios: Tensor,  # [N]
    binoms: Tensor,  # [n_max, n_max]
) -> Tuple[Tensor, Tensor]:
    """New opacities [N] and scales [N,3] of Gaussians that are split into `ratios` copies each (Eq. 9 of
    "3D Gaussian Splatting as Markov Chain Monte Carlo"); `binoms` is the lookup table of binomial coefficients.
    As in the reference, `ratios` is clamped in place to [1, n_max]."""
    N = opacities.shape[0]
    n_max, _ = binoms.shape
    assert scales.shape == (N, 3), scales.shape
    assert ratios.shape == (N,), ratios.shape
    for t in (opacities, scales, binoms):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError("compute_relocation: float32 CUDA tensors expected (no CPU implementation)")
    opacities = opacities.contiguous()
    scales = scales.contiguous()
    binoms = binoms.contiguous()
    ratios.clamp_(min=1, max=n_max)
    ratios = ratios.int().contiguous()
    new_opacities = torch.empty_like(opacities)
    new_scales = torch.empty_like(scales)
    with torch.cuda.device(opacities.device):
        call("gsb200_compute_relocation", N, ptr(opacities), ptr(scales), ptr(ratios), ptr(binoms), n_max, ptr(new_opacities),
             ptr(new_scales), stream_ptr(opacities.device))
    return new_opacities, new_scales
