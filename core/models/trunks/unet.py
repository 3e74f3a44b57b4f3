"""Shim for the reference's core/models/trunks/unet.py -> im2im_uq_b200.models.unet."""
from im2im_uq_b200.models.unet import DoubleConv, Down, OutConv, UNet, Up  # noqa: F401
