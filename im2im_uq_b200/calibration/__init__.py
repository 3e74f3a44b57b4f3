"""RCPS calibration entry points with the reference's names (core/calibration in aangelopoulos/im2im-uq)."""
