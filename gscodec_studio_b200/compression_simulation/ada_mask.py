"""Annealed learnable mask for the higher-order SH coefficients (reference: compression_simulation/
ada_mask.py:6-60).  Elementwise torch on [N,1,1] logits -- not a kernel of the hot path."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def annealed_temperature(step: int, start_iter: int, total_iters: int, start_temp: float, end_temp: float) -> float:
    """Constant `start_temp` until `start_iter`, then a geometric decay that reaches `end_temp` at `total_iters`."""
    if step < start_iter:
        return start_temp
    progress = min(max((step - start_iter) / (total_iters - start_iter), 0), 1)
    return start_temp * (end_temp / start_temp) ** progress


class AnnealingMask(nn.Module):
    """x * sigmoid(logits / T(step)) while training (T annealed from `start_temp` to `end_temp`), x * [sigmoid(logit) >= 0.5]
    in eval mode; `get_sparsity_loss` pulls the mean soft mask towards `target_sparsity`."""

    def __init__(self, input_shape, device, total_iters=30_000, start_temp=5.0, end_temp=0.1,
                 annealing_start_iter=10_000, target_sparsity=0.2):
        super().__init__()
        self.mask_logits = nn.Parameter(torch.ones(input_shape, device=device))
        self.total_iters, self.annealing_start_iter = total_iters, annealing_start_iter
        self.start_temp, self.end_temp = start_temp, end_temp
        self.target_sparsity = target_sparsity
        self.current_iter = 0

    def get_temperature(self, current_step):
        return annealed_temperature(current_step, self.annealing_start_iter, self.total_iters, self.start_temp, self.end_temp)

    def _soft_mask(self, step):
        return torch.sigmoid(self.mask_logits / self.get_temperature(step))

    def forward(self, x, current_step):
        if not self.training:
            return x * self.get_binary_mask()
        self.current_iter = current_step
        return x * self._soft_mask(current_step)

    @torch.no_grad()
    def get_binary_mask(self):
        return (torch.sigmoid(self.mask_logits) >= 0.5).float()

    def get_sparsity_loss(self, lambda_l1=0.01, lambda_target=0.1):
        mean_mask = self._soft_mask(self.current_iter).mean()
        target = torch.full_like(mean_mask, self.target_sparsity)
        return lambda_l1 * mean_mask + lambda_target * F.binary_cross_entropy(mean_mask, target)

    @torch.no_grad()
    def get_mask_ratio(self):
        return self.get_binary_mask().sum() / self.mask_logits.shape[0]
