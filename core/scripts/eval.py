"""Shim for the reference's core/scripts/eval.py -> im2im_uq_b200.scripts.eval (the calibration-side entry points)."""
from im2im_uq_b200.scripts.eval import eval_set_metrics, get_loss_table  # noqa: F401
