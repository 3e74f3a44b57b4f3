"""Shim for the reference's core/calibration/calibrate_model.py -> im2im_uq_b200.calibration.calibrate_model."""
from im2im_uq_b200.calibration.calibrate_model import *  # noqa: F401,F403
from im2im_uq_b200.calibration.calibrate_model import (calibrate_model, calibrate_from_outputs, evaluate_from_loss_table,  # noqa: F401
                                                       fraction_missed_loss, get_rcps_loss_fn,
                                                       get_rcps_losses_from_outputs, get_rcps_metrics_from_outputs)
