"""MCMC relocation op -- the API of the reference's ``gsplat/relocation.py:10-49`` on the libgsb200 kernel."""
from typing import Tuple

import torch
from torch import Tensor

from ._cabi import call, device_guard, ptr, stream_ptr


def compute_relocation(opacities: Tensor, scales: Tensor, ratios: Tensor, binoms: Tensor) -> Tuple[Tensor, Tensor]:
    """Opacity and scale of Gaussians that are each split into ``ratios`` copies (Eq. 9 of "3D Gaussian Splatting as
    Markov Chain Monte Carlo", arXiv:2404.09591).

    opacities [N], scales [N,3], ratios [N] (any integer-valued dtype), binoms [n_max, n_max] = table of binomial
    coefficients.  Returns (new_opacities [N], new_scales [N,3]).  As in the reference, ``ratios`` is clamped IN
    PLACE to [1, n_max] before use."""
    n_gauss, n_max = opacities.shape[0], binoms.shape[0]
    assert scales.shape == (n_gauss, 3), scales.shape
    assert ratios.shape == (n_gauss,), ratios.shape
    if any((not t.is_cuda) or t.dtype != torch.float32 for t in (opacities, scales, binoms)):
        raise RuntimeError("compute_relocation: float32 CUDA tensors expected (no CPU implementation)")
    ratios.clamp_(min=1, max=n_max)
    opac_c, scales_c, binoms_c, ratios_i = opacities.contiguous(), scales.contiguous(), binoms.contiguous(), ratios.int().contiguous()
    out_opacities, out_scales = torch.empty_like(opac_c), torch.empty_like(scales_c)
    with device_guard(opacities.device):
        call("gsb200_compute_relocation", n_gauss, ptr(opac_c), ptr(scales_c), ptr(ratios_i), ptr(binoms_c), n_max,
             ptr(out_opacities), ptr(out_scales), stream_ptr(opacities.device))
    return out_opacities, out_scales
