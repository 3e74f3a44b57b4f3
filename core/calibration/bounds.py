"""Shim for the reference's core/calibration/bounds.py -> im2im_uq_b200.calibration.bounds."""
from im2im_uq_b200.calibration.bounds import HB_mu_plus, bentkus_plus, h1, hoeffding_plus  # noqa: F401
