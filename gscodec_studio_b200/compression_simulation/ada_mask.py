"""Annealed learnable mask for the higher-order SH coefficients (reference: compression_simulation/
ada_mask.py:6-60).  Elementwise torch on [N,1,1] logits -- not a kernel of the hot path."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class AnnealingMask(nn.Module):
    def __init__(self, input_shape, device, total_iters=30_000, start_temp=5.0, end_temp=0.1,
                 annealing_start_iter=10_000, target_sparsity=0.2):
        super().__init__()
        self.mask_logits = nn.Parameter(torch.ones(input_shape, device=device))
        self.total_iters = total_iters
        self.start_temp = start_temp
        self.end_temp = end_temp
        self.annealing_start_iter = annealing_start_iter
        self.current_iter = 0
        self.target_sparsity = target_sparsity

    def get_temperature(self, current_step):
        if current_step < self.annealing_start_iter:
            return self.start_temp
        span = self.total_iters - self.annealing_start_iter
        progress = min(max((current_step - self.annealing_start_iter) / span, 0), 1)
        return self.start_temp * math.exp(math.log(self.end_temp / self.start_temp) * progress)

    def _soft_mask(self, step):
        return torch.sigmoid(self.mask_logits / self.get_temperature(step))

    def forward(self, x, current_step):
        if self.training:
            self.current_iter = current_step
            return x * self._soft_mask(current_step)
        return x * self.get_binary_mask()

    @torch.no_grad()
    def get_binary_mask(self):
        return (torch.sigmoid(self.mask_logits) >= 0.5).float()

    def get_sparsity_loss(self, lambda_l1=0.01, lambda_target=0.1):
        mean_mask = torch.mean(self._soft_mask(self.current_iter))
        target = torch.tensor(self.target_sparsity).to(mean_mask.device)
        return lambda_l1 * mean_mask + lambda_target * F.binary_cross_entropy(mean_mask, target)

    @torch.no_grad()
    def get_mask_ratio(self):
        return torch.sum(torch.sigmoid(self.mask_logits) >= 0.5) / self.mask_logits.shape[0]
