"""Shim for the reference's core/models/losses/pinball.py -> im2im_uq_b200.models.pinball."""
from im2im_uq_b200.models.pinball import PinballLoss  # noqa: F401
