"""Shim for the reference's core/models/finallayers/quantile_l1_layer.py -> im2im_uq_b200.models.heads."""
from im2im_uq_b200.models.heads import (  # noqa: F401
    QuantileRegressionL1Layer, quantile_regression_l1_loss_fn, quantile_regression_l1_nested_sets_from_output)
