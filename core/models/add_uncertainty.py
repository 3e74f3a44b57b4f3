"""Shim for the reference's core/models/add_uncertainty.py -> im2im_uq_b200.models.add_uncertainty."""
from im2im_uq_b200.models.add_uncertainty import ModelWithUncertainty, add_uncertainty  # noqa: F401
