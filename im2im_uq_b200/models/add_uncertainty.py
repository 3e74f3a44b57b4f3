"""The reference's plugin surface - mirror of ``core/models/add_uncertainty.py``.

  ModelWithUncertainty   :15-49   forward / loss_fn / nested_sets_from_output / nested_sets / set_lhat, buffer lhat
  add_uncertainty        :51-87   trunk + head selected by params["uncertainty_type"]

All seven ``uncertainty_type`` values of the reference are available; every head's set function and calibration run
on the native kernels (heads.py), the quantile head also has the native UNet inference/training engines behind
``forward``; the other heads' forwards use the engine for the trunk + stacked head convolution where it applies.
"""
import torch
import torch.nn as nn

from . import heads
from .quantile_layer import (QuantileRegressionLayer, quantile_regression_loss_fn,
                             quantile_regression_nested_sets_from_output)


class ModelWithUncertainty(nn.Module):
    def __init__(self, baseModel, last_layer, in_train_loss_fn, in_nested_sets_from_output_fn, params):
        super(ModelWithUncertainty, self).__init__()
        self.baseModel = baseModel
        self.last_layer = last_layer
        self.register_buffer('lhat', None)
        self.in_train_loss_fn = in_train_loss_fn
        self.in_nested_sets_from_output_fn = in_nested_sets_from_output_fn
        self.params = params

    def forward(self, x):
        # eval-mode CUDA forwards under no_grad run on the native sm_100a engine (tcgen05 convolutions);
        # training forwards go through the module graph so autograd sees them.
        from .unet_engine import UNetInferenceEngine, native_forward_applicable
        if native_forward_applicable(self, x):
            eng = self.__dict__.get("_native_engine")
            if eng is None:
                eng = UNetInferenceEngine(self)
                self.__dict__["_native_engine"] = eng
            return eng.forward(x)
        from .unet_train import native_train_applicable, native_train_forward
        if native_train_applicable(self, x):
            return native_train_forward(self, x)  # training step on the native engine, bridged into autograd
        if torch.is_tensor(x) and x.is_cuda and not self.__dict__.get("_warned_module_path"):
            # CUDA input that neither engine takes (eval mode with autograd on, a trunk that is not the UNet, a disabled
            # engine ...): say so once instead of silently running the library convolutions
            import warnings
            self.__dict__["_warned_module_path"] = True
            warnings.warn("ModelWithUncertainty.forward: this call runs through the torch module graph, not the native "
                          "sm_100a engines (they need the reference's bilinear UNet trunk and either eval mode under "
                          "torch.no_grad() or training mode with autograd enabled)", stacklevel=2)
        x = self.baseModel(x)
        return self.last_layer(x)

    def loss_fn(self, pred, target):
        return self.in_train_loss_fn(pred, target, self.params)

    # Always outputs [0,1] valued nested sets
    def nested_sets_from_output(self, output, lam=None):
        lower_edge, prediction, upper_edge = self.in_nested_sets_from_output_fn(self, output, lam)
        if getattr(self.in_nested_sets_from_output_fn, "im2im_head_kind", None) is None:
            # user-supplied set functions: apply the reference's lower bound on the set size (:35-36); the built-in
            # heads have it fused into their kernel (the clamp is idempotent)
            upper_edge = torch.maximum(upper_edge, prediction + 1e-6)
            lower_edge = torch.minimum(lower_edge, prediction - 1e-6)
        return lower_edge, prediction, upper_edge

    def nested_sets(self, x, lam=None):
        if lam is None:
            if self.lhat is None:
                raise Exception("You have to specify lambda unless your model is already calibrated.")
            lam = self.lhat
        output = self(*x)
        return self.nested_sets_from_output(output, lam=lam)

    def set_lhat(self, lhat):
        self.lhat = lhat


_OTHER_HEADS = {
    "quantiles_l1": (heads.QuantileRegressionL1Layer, heads.quantile_regression_l1_loss_fn,
                     heads.quantile_regression_l1_nested_sets_from_output),
    "gaussian": (heads.GaussianRegressionLayer, heads.gaussian_regression_loss_fn,
                 heads.gaussian_regression_nested_sets_from_output),
    "residual_magnitude": (heads.ResidualMagnitudeLayer, heads.residual_magnitude_loss_fn,
                           heads.residual_magnitude_nested_sets_from_output),
    "residual_magnitude_l1": (heads.ResidualMagnitudeL1Layer, heads.residual_magnitude_l1_loss_fn,
                              heads.residual_magnitude_l1_nested_sets_from_output),
    "softmax": (heads.SoftmaxLayer, heads.softmax_loss_fn, heads.softmax_nested_sets_from_output),
    "inn": (heads.INNLayer, heads.inn_loss_fn, heads.inn_nested_sets_from_output),
}


def add_uncertainty(model, params):
    if params["uncertainty_type"] == "quantiles":
        last_layer = QuantileRegressionLayer(model.n_channels_middle, model.n_channels_out, params)
        train_loss_fn = quantile_regression_loss_fn
        nested_sets_from_output_fn = quantile_regression_nested_sets_from_output
    elif params["uncertainty_type"] in _OTHER_HEADS:
        layer_cls, train_loss_fn, nested_sets_from_output_fn = _OTHER_HEADS[params["uncertainty_type"]]
        last_layer = layer_cls(model.n_channels_middle, model.n_channels_out, params)
    else:
        raise NotImplementedError
    return ModelWithUncertainty(model, last_layer, train_loss_fn, nested_sets_from_output_fn, params)
