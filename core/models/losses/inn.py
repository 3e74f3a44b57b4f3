"""Shim for the reference's core/models/losses/inn.py -> im2im_uq_b200.models.inn."""
from im2im_uq_b200.models.inn import INNLoss  # noqa: F401
