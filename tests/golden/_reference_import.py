"""Import the UNMODIFIED reference (aangelopoulos/im2im-uq) from /root/reference for fixture generation.

Only usable in the authoring container (the GPU box has no /root/reference).  The reference's
``core/utils.py:5-7`` imports matplotlib at module top level, which is not installed here; empty stub
modules are pre-inserted so that ``core.models.add_uncertainty`` imports.  Nothing in the
reference tree is modified or copied.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("IM2IM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core", "calibration"))


def import_reference():
    """Returns a namespace with the reference modules used as the live oracle."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the mount is read-only
    os.environ.setdefault("WANDB_MODE", "disabled")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    # our own drop-in shim is also called `core`; make sure the reference's package wins in this process
    for name in [m for m in sys.modules if m == "core" or m.startswith("core.")]:
        del sys.modules[name]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    import core.calibration.bounds as bounds
    import core.calibration.calibrate_model as calibrate_model
    import core.models.add_uncertainty as add_uncertainty
    import core.models.finallayers.quantile_layer as quantile_layer
    import core.models.finallayers.gaussian_layer as gaussian_layer
    import core.models.finallayers.inn_layer as inn_layer
    import core.models.finallayers.quantile_l1_layer as quantile_l1_layer
    import core.models.finallayers.residual_magnitude_l1_layer as residual_magnitude_l1_layer
    import core.models.finallayers.residual_magnitude_layer as residual_magnitude_layer
    import core.models.finallayers.softmax_layer as softmax_layer
    import core.models.losses.inn as inn
    import core.models.losses.pinball as pinball
    import core.models.trunks.unet as unet
    assert bounds.__file__.startswith(REFERENCE_ROOT), bounds.__file__
    ns = types.SimpleNamespace(bounds=bounds, calibrate_model=calibrate_model, add_uncertainty=add_uncertainty,
                               quantile_layer=quantile_layer, pinball=pinball, unet=unet, gaussian_layer=gaussian_layer,
                               inn_layer=inn_layer, quantile_l1_layer=quantile_l1_layer,
                               residual_magnitude_layer=residual_magnitude_layer,
                               residual_magnitude_l1_layer=residual_magnitude_l1_layer, softmax_layer=softmax_layer,
                               inn=inn)
    return ns
