"""Shim for the reference's core/models/finallayers/gaussian_layer.py -> im2im_uq_b200.models.heads."""
from im2im_uq_b200.models.heads import (  # noqa: F401
    GaussianRegressionLayer, gaussian_regression_loss_fn, gaussian_regression_nested_sets_from_output)
