"""Shim for the reference's core/models/finallayers/quantile_layer.py -> im2im_uq_b200.models.quantile_layer."""
from im2im_uq_b200.models.quantile_layer import (QuantileRegressionLayer, quantile_regression_loss_fn,  # noqa: F401
                                                 quantile_regression_nested_sets_from_output)
