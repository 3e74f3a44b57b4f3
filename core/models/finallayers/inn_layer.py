"""Shim for the reference's core/models/finallayers/inn_layer.py -> im2im_uq_b200.models.heads."""
from im2im_uq_b200.models.heads import (  # noqa: F401
    INNLayer, inn_loss_fn, inn_nested_sets_from_output)
