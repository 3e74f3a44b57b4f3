"""Quantile-regression head - mirror of the reference's ``core/models/finallayers/quantile_layer.py``.

  QuantileRegressionLayer                      :8-21   three 3x3 convs middle->out stacked on a new dim 1
  quantile_regression_loss_fn                  :23-32  w_lo*pinball(q_lo) + w_hi*pinball(q_hi) + w_mse*MSE
  quantile_regression_nested_sets_from_output  :34-44  clamp + lam-scaled interval (CUDA kernel, fp32 op order kept)
"""
import torch
import torch.nn as nn

from .. import rcps
from .pinball import PinballLoss


class QuantileRegressionLayer(nn.Module):
    def __init__(self, n_channels_middle, n_channels_out, params):
        super(QuantileRegressionLayer, self).__init__()
        self.q_lo = params["q_lo"]
        self.q_hi = params["q_hi"]
        self.params = params
        # same parameter names/shapes/creation order as the reference, so checkpoints and seeded inits line up
        self.lower = nn.Conv2d(n_channels_middle, n_channels_out, kernel_size=3, padding=1)
        self.prediction = nn.Conv2d(n_channels_middle, n_channels_out, kernel_size=3, padding=1)
        self.upper = nn.Conv2d(n_channels_middle, n_channels_out, kernel_size=3, padding=1)

    def forward(self, x):
        # one conv with the three heads stacked on the output-channel axis == three convs + unsqueeze/cat
        w = torch.cat((self.lower.weight, self.prediction.weight, self.upper.weight), dim=0)
        b = torch.cat((self.lower.bias, self.prediction.bias, self.upper.bias), dim=0)
        y = nn.functional.conv2d(x, w, b, padding=1)
        n, _, h, wd = y.shape
        return y.view(n, 3, -1, h, wd)


def quantile_regression_loss_fn(pred, target, params):
    if (pred.is_cuda and pred.dtype == torch.float32 and pred.dim() == 5 and pred.shape[1] == 3
            and target.numel() * 3 == pred.numel()):
        from .unet_train import native_quantile_loss
        return native_quantile_loss(pred, target, params)  # fused pinball+MSE forward/backward kernel
    q_lo_loss = PinballLoss(quantile=params["q_lo"])
    q_hi_loss = PinballLoss(quantile=params["q_hi"])
    mse_loss = nn.MSELoss()
    loss = params['q_lo_weight'] * q_lo_loss(pred[:, 0, :, :, :].squeeze(), target.squeeze()) + \
        params['q_hi_weight'] * q_hi_loss(pred[:, 2, :, :, :].squeeze(), target.squeeze()) + \
        params['mse_weight'] * mse_loss(pred[:, 1, :, :, :].squeeze(), target.squeeze())
    return loss


def quantile_regression_nested_sets_from_output(model, output, lam=None):
    """(lower_edge, prediction, upper_edge) at ``lam`` (default: the calibrated ``model.lhat``).

    Like the reference this clamps ``output[:,0]``/``output[:,2]`` in place (:39-40) and returns ``output[:,1]`` as a
    view.  The +/-1e-6 clamp that ``ModelWithUncertainty.nested_sets_from_output`` applies on top
    (add_uncertainty.py:35-36) is idempotent on these values, so the fused kernel result is final.
    """
    if lam is None:
        if model.lhat is None:
            raise Exception("You have to specify lambda unless your model is already calibrated.")
        lam = model.lhat
    lower_edge, prediction, upper_edge = rcps.quantile_nested_sets(output, float(lam), write_back_clamp=True)
    return lower_edge, prediction, upper_edge


quantile_regression_nested_sets_from_output.im2im_head_kind = 0        # _lib.IM2IM_HEAD_QUANTILES
quantile_regression_nested_sets_from_output.im2im_scores_from_output = None
