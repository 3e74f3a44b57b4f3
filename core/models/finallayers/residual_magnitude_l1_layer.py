"""Shim for the reference's core/models/finallayers/residual_magnitude_l1_layer.py -> im2im_uq_b200.models.heads."""
from im2im_uq_b200.models.heads import (  # noqa: F401
    ResidualMagnitudeL1Layer, residual_magnitude_l1_loss_fn, residual_magnitude_l1_nested_sets_from_output)
