"""Interval-score style loss of the INN head - mirror of the reference's ``core/models/losses/inn.py`` (INNLoss :4-22):
squared overshoot of the target above ``upper`` and below ``lower`` plus ``beta`` times the interval width."""
import torch


class INNLoss():
    def __init__(self, beta=0.10, reduction='mean'):
        assert 0 <= beta
        self.beta, self.reduction = beta, reduction

    def __call__(self, lower, upper, target):
        assert target.shape == lower.shape and target.shape == upper.shape
        above = torch.clamp_min(target - upper, 0)
        below = torch.clamp_min(lower - target, 0)
        loss = above * above + below * below + self.beta * (upper - lower).abs()
        if self.reduction == 'sum':
            return loss.sum()
        if self.reduction == 'mean':
            return loss.mean()
        return loss
