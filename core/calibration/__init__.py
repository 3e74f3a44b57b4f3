"""core.calibration shim -> im2im_uq_b200.calibration"""
