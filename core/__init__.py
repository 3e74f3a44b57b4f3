"""Drop-in shim: the reference's import paths (``core.calibration...``, ``core.models...``) backed by im2im_uq_b200.

A caller written against aangelopoulos/im2im-uq (e.g. its core/scripts/router.py) keeps its imports and gets the
B200 hot path.  Nothing here is copied from the reference; every module re-exports the package's implementation.
"""
