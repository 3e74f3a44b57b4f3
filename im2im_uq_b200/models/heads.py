"""The reference's other uncertainty heads on the native kernels.

One module mirrors six reference files (``core/models/finallayers/``); ``core/`` re-exports the names under the
reference's import paths:

  uncertainty_type         layer (state-dict names)                         set function / training loss
  "gaussian"               GaussianRegressionLayer (mean, variance)         gaussian_layer.py:26-34 / :20-24
  "residual_magnitude"     ResidualMagnitudeLayer (prediction,              residual_magnitude_layer.py:28-36 / :20-26
                           residual_magnitude)
  "residual_magnitude_l1"  ResidualMagnitudeL1Layer (same)                  residual_magnitude_l1_layer.py:28-36 / :20-26
  "quantiles_l1"           QuantileRegressionL1Layer (lower, prediction,    quantile_l1_layer.py:34-44 / :23-32
                           upper)
  "inn"                    INNLayer (lower, prediction, upper)              inn_layer.py:30-40 / :23-28, losses/inn.py
  "softmax"                SoftmaxLayer (output_layers[i])                  softmax_layer.py:27-53 / :16-25

What runs where: every set function is ONE launch of ``im2im_nested_sets`` (head kind selects the width formula; the
+/-1e-6 clamp of add_uncertainty.py:35-36 is fused, so the values returned here are already the final ones - the clamp
is idempotent); the calibration sweep uses the same head kinds in ``im2im_rcps_miss_counts``; the training losses are
``im2im_head_loss_f32`` (value + gradient in one pass).  The softmax head's lambda-independent half (softmax, cumsum,
quantile counts, argmax) is ``im2im_softmax_sets`` and is computed once per calibration instead of once per lambda.
CPU tensors take plain torch formulas in the loss functions only (module construction / unit tests); set functions and
calibration have no CPU path.
"""
import torch
import torch.nn as nn

from .. import _lib, rcps
from .pinball import PinballLoss


def _resolve_lam(model, lam):
    if lam is None:
        if model.lhat is None:
            raise Exception("You have to specify lambda unless your model is already calibrated.")
        lam = model.lhat
    return float(lam)


def _conv(c_in, c_out):
    return nn.Conv2d(c_in, c_out, kernel_size=3, padding=1)


class _StackedPlanesHead(nn.Module):
    """n planes, each a 3x3 conv middle->out, stacked on a new dim 1; planes >= act_from get `act` (relu / abs).

    Subclasses create their convolutions under the reference's attribute names, in the reference's order (seeded
    initialisations and checkpoints line up), and list them in ``plane_names``."""
    plane_names = ()
    act = None          # None | "relu" | "abs"
    act_from = 0        # first plane (in units of planes) the activation applies to

    def plane_convs(self):
        return [getattr(self, n) for n in self.plane_names]

    def forward(self, x):
        convs = self.plane_convs()
        y = nn.functional.conv2d(x, torch.cat([c.weight for c in convs], 0), torch.cat([c.bias for c in convs], 0),
                                 padding=1)
        n, _, h, w = y.shape
        y = y.view(n, len(convs), -1, h, w)
        if self.act is not None:
            head, tail = y[:, :self.act_from], y[:, self.act_from:]
            y = torch.cat((head, torch.relu(tail) if self.act == "relu" else tail.abs()), dim=1)
        return y


class GaussianRegressionLayer(_StackedPlanesHead):
    plane_names, act, act_from = ("mean", "variance"), "relu", 1

    def __init__(self, n_channels_middle, n_channels_out, params):
        super().__init__()
        self.params = params
        self.mean = _conv(n_channels_middle, n_channels_out)
        self.variance = _conv(n_channels_middle, n_channels_out)


class ResidualMagnitudeLayer(_StackedPlanesHead):
    plane_names, act, act_from = ("prediction", "residual_magnitude"), "abs", 1

    def __init__(self, n_channels_middle, n_channels_out, params):
        super().__init__()
        self.params = params
        self.prediction = _conv(n_channels_middle, n_channels_out)
        self.residual_magnitude = _conv(n_channels_middle, n_channels_out)


class ResidualMagnitudeL1Layer(ResidualMagnitudeLayer):
    pass


class QuantileRegressionL1Layer(_StackedPlanesHead):
    plane_names = ("lower", "prediction", "upper")

    def __init__(self, n_channels_middle, n_channels_out, params):
        super().__init__()
        self.q_lo, self.q_hi, self.params = params["q_lo"], params["q_hi"], params
        self.lower = _conv(n_channels_middle, n_channels_out)
        self.prediction = _conv(n_channels_middle, n_channels_out)
        self.upper = _conv(n_channels_middle, n_channels_out)


class INNLayer(_StackedPlanesHead):
    plane_names = ("lower", "prediction", "upper")

    def __init__(self, n_channels_middle, n_channels_out, params):
        super().__init__()
        self.beta, self.params = params["beta"], params
        self.lower = _conv(n_channels_middle, n_channels_out)
        self.prediction = _conv(n_channels_middle, n_channels_out)
        self.upper = _conv(n_channels_middle, n_channels_out)


class SoftmaxLayer(nn.Module):
    """n_channels_out classifiers over num_softmax bins; output (B, num_softmax*n_channels_out, 1, H, W)."""

    def __init__(self, n_channels_middle, n_channels_out, params):
        super().__init__()
        self.num_softmax = params["num_softmax"]
        self.output_layers = nn.ModuleList([_conv(n_channels_middle, self.num_softmax) for _ in range(n_channels_out)])

    def forward(self, x):
        w = torch.cat([layer.weight for layer in self.output_layers], 0)
        b = torch.cat([layer.bias for layer in self.output_layers], 0)
        return nn.functional.conv2d(x, w, b, padding=1).unsqueeze(2)


# ------------------------------------------------------------------------------------------------ training losses
class _HeadLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, kind, q_lo, q_hi, w0, w1, w2, beta):
        lib = _lib.load()
        pred_c, target_c = pred.contiguous(), target.contiguous().float()
        n, px = pred_c.shape[0], target_c[0].numel()
        dpred = torch.empty_like(pred_c)
        parts = torch.empty(3, dtype=torch.float64, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(lib.im2im_head_loss_f32(kind, pred_c.data_ptr(), target_c.data_ptr(), n, px, q_lo, q_hi, w0, w1, w2,
                                               beta, dpred.data_ptr(), parts.data_ptr(),
                                               torch.cuda.current_stream(pred.device).cuda_stream), "im2im_head_loss_f32")
        ctx.save_for_backward(dpred)
        return ((parts[0] * w0 + parts[1] * w1 + parts[2] * w2) / float(n * px)).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return (dpred * g,) + (None,) * 8


def _native_loss_ok(pred, target, planes):
    return (pred.is_cuda and pred.dtype == torch.float32 and pred.dim() == 5 and pred.shape[1] == planes
            and target.numel() * planes == pred.numel())


def _native_loss(kind, pred, target, q_lo=0.5, q_hi=0.5, w=(1.0, 1.0, 1.0), beta=0.0):
    return _HeadLossFn.apply(pred, target, kind, float(q_lo), float(q_hi), float(w[0]), float(w[1]), float(w[2]),
                             float(beta))


def gaussian_regression_loss_fn(pred, target, params):
    if _native_loss_ok(pred, target, 2):
        # nn.GaussianNLLLoss raises on a negative variance; the layer's relu cannot produce one, and checking here would
        # cost a device->host sync per training step, so the fused kernel just applies the eps clamp
        return _native_loss(_lib.LOSS_GAUSSIAN, pred, target)
    return nn.GaussianNLLLoss()(pred[:, 0].squeeze(), target.squeeze(), pred[:, 1].squeeze())


def _residual_loss(pred, target, l1: bool):
    if _native_loss_ok(pred, target, 2):
        return _native_loss(_lib.LOSS_RESIDUAL_L1 if l1 else _lib.LOSS_RESIDUAL, pred, target)
    centre, mag, y = pred[:, 0].squeeze(), pred[:, 1].squeeze(), target.squeeze()
    first = nn.functional.l1_loss(centre, y) if l1 else nn.functional.mse_loss(centre, y)
    return first + nn.functional.mse_loss(mag, (y - centre).abs())


def residual_magnitude_loss_fn(pred, target, params):
    return _residual_loss(pred, target, l1=False)


def residual_magnitude_l1_loss_fn(pred, target, params):
    return _residual_loss(pred, target, l1=True)


def quantile_regression_l1_loss_fn(pred, target, params):
    w = (params['q_lo_weight'], params['q_hi_weight'], params['mse_weight'])
    if _native_loss_ok(pred, target, 3):
        return _native_loss(_lib.LOSS_QUANTILES_L1, pred, target, params["q_lo"], params["q_hi"], w)
    y = target.squeeze()
    return (w[0] * PinballLoss(quantile=params["q_lo"])(pred[:, 0].squeeze(), y)
            + w[1] * PinballLoss(quantile=params["q_hi"])(pred[:, 2].squeeze(), y)
            + w[2] * nn.functional.l1_loss(pred[:, 1].squeeze(), y))


def inn_loss_fn(pred, target, params):
    if _native_loss_ok(pred, target, 3):
        return _native_loss(_lib.LOSS_INN, pred, target, beta=params["beta"])
    from .inn import INNLoss
    y = target.squeeze()
    return nn.functional.mse_loss(pred[:, 1].squeeze(), y) + INNLoss(beta=params["beta"])(pred[:, 0].squeeze(),
                                                                                          pred[:, 2].squeeze(), y)


def softmax_loss_fn(pred, target, params):
    """Cross entropy against the label's bin (softmax_layer.py:16-25); library op, not on the accelerated path."""
    classes = torch.linspace(0, 1, params["num_softmax"], device=pred.device)
    idx = torch.bucketize(target, classes, right=False).clamp_(max=params["num_softmax"] - 1)
    return nn.functional.cross_entropy(pred, idx)


# ------------------------------------------------------------------------------------------------ set functions
def _set_fn(head_kind, scores_from_output=None, doc=""):
    def fn(model, output, lam=None):
        lam = _resolve_lam(model, lam)
        with torch.no_grad():
            scores = scores_from_output(output) if scores_from_output is not None else output
            return rcps.head_nested_sets(scores, lam, head_kind)
    fn.__doc__ = doc
    fn.im2im_head_kind = head_kind                  # the calibration sweep runs the same head kind in one pass
    fn.im2im_scores_from_output = scores_from_output
    return fn


gaussian_regression_nested_sets_from_output = _set_fn(
    _lib.IM2IM_HEAD_GAUSSIAN, doc="mean -/+ lam*sqrt(variance) (gaussian_layer.py:26-34), clamp fused")
residual_magnitude_nested_sets_from_output = _set_fn(
    _lib.IM2IM_HEAD_RESIDUAL, doc="pred -/+ lam*|residual| (residual_magnitude_layer.py:28-36), clamp fused")
residual_magnitude_l1_nested_sets_from_output = _set_fn(
    _lib.IM2IM_HEAD_RESIDUAL, doc="pred -/+ lam*|residual| (residual_magnitude_l1_layer.py:28-36), clamp fused")
softmax_nested_sets_from_output = _set_fn(
    _lib.IM2IM_HEAD_SOFTMAX_SETS, scores_from_output=rcps.softmax_sets,
    doc="argmax -/+ lam*relu(distance to the 5%/95% softmax quantile) (softmax_layer.py:27-53), clamp fused")


def _quantile_type_set_fn(doc):
    from .quantile_layer import quantile_regression_nested_sets_from_output as base

    def fn(model, output, lam=None):
        return base(model, output, lam)   # same arithmetic incl. the in-place clamp of `output` (:39-40)
    fn.__doc__ = doc
    fn.im2im_head_kind = _lib.IM2IM_HEAD_QUANTILES
    fn.im2im_scores_from_output = None
    return fn


quantile_regression_l1_nested_sets_from_output = _quantile_type_set_fn("quantile_l1_layer.py:34-44")
inn_nested_sets_from_output = _quantile_type_set_fn("inn_layer.py:30-40")
