"""Shim for the reference's core/models/finallayers/softmax_layer.py -> im2im_uq_b200.models.heads."""
from im2im_uq_b200.models.heads import (  # noqa: F401
    SoftmaxLayer, softmax_loss_fn, softmax_nested_sets_from_output)
