"""Pinball (quantile) loss - mirror of the reference's ``core/models/losses/pinball.py`` (PinballLoss :4-26).

The reference builds the loss with boolean-mask assignment (``loss[mask] = ...``), which forces a ``nonzero`` and a
host sync per call; the closed form  q*|e| for e<0, (1-q)*|e| for e>0, 0 for e==0  (e = output - target) is the same
value element for element and is what is evaluated here.
"""
import torch


class PinballLoss():
    def __init__(self, quantile=0.10, reduction='mean'):
        self.quantile = quantile
        assert 0 < self.quantile
        assert self.quantile < 1
        self.reduction = reduction

    def __call__(self, output, target):
        assert output.shape == target.shape
        error = output - target
        mag = error.abs()
        # identical products to the reference: quantile*|e| where e<0, (1-quantile)*|e| where e>0
        loss = torch.where(error < 0, self.quantile * mag, torch.where(error > 0, (1 - self.quantile) * mag,
                                                                    torch.zeros_like(mag)))
        if self.reduction == 'sum':
            loss = loss.sum()
        if self.reduction == 'mean':
            loss = loss.mean()
        return loss
