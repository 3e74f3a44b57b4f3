"""CPU: the C-ABI library loads and exports every symbol the header declares; the product never falls back to CPU."""
import os
import re

import pytest
import torch

from conftest import ROOT
from im2im_uq_b200 import _lib


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "im2im_uq.h")).read()
    declared = set(re.findall(r"\b(im2im_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/im2im_uq.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert lib.im2im_abi_version() == 1


def test_library_has_sm100a_code_and_no_torch_dependency():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_argument_validation_without_gpu():
    lib = _lib.load()
    rc = lib.im2im_rcps_miss_counts(None, None, None, None, 1, 16, 16, 16, 16, 16, None, 0, 0, None, None, 0, None)
    assert rc == -34 and b"n_lambdas" in lib.im2im_last_error()
    rc = lib.im2im_rcps_miss_counts(None, None, None, None, 1, 16, 16, 16, 16, 16, None, 10, 7, None, None, 0, None)
    assert rc == -95
    rc = lib.im2im_rcps_miss_counts(None, None, None, None, 1, 1 << 24, 0, 0, 0, 0, None, 10, 0, None, None, 0, None)
    assert rc == -34
    rc = lib.im2im_rcps_miss_counts(None, None, None, None, -1, 16, 16, 16, 16, 16, None, 10, 0, None, None, 0, None)
    assert rc == -22
    with pytest.raises(_lib.Im2ImError):
        _lib.check(rc, "probe")
    # entry points added for the other heads / the graphed training step: argument checks come before any CUDA call
    assert lib.im2im_nested_sets(9, None, None, None, 1, 16, 16, 16, 16, 1.0, 0, None, None, None) == -95
    assert lib.im2im_nested_sets(2, None, None, None, 1, 16, 16, 16, 16, 1.0, 0, None, None, None) == -22
    assert lib.im2im_nested_sets(1, None, None, None, 0, 16, 16, 16, 16, 1.0, 0, None, None, None) == 0   # empty: no-op
    assert lib.im2im_rcps_miss_map(None, None, None, None, 1, 16, 16, 16, 16, 16, 1.0, 5, None, 0, None) == -95
    assert lib.im2im_softmax_sets(None, 1, 65, 16, 16, 16, None, None) == -34 and b"n_classes" in lib.im2im_last_error()
    assert lib.im2im_softmax_sets(None, 1, 50, 16, 800, 16, None, None) == -22
    assert lib.im2im_head_loss_f32(2, None, None, 1, 16, .05, .95, 1., 1., 1., .1, None, None, None) == -22
    assert lib.im2im_head_conv3x3_act_f32(None, None, None, None, 1, 8, 8, 32, 32, 2, 7, 1, None, None) == -22
    assert lib.im2im_adam_step_dev_f32(None, None, None, None, 10, 1e-3, .9, .999, 1e-8, None, 1.0, None) == -22


def test_cpu_tensors_are_rejected_loudly():
    from im2im_uq_b200 import rcps
    from im2im_uq_b200.calibration import calibrate_model as cm
    out = torch.zeros(2, 3, 1, 4, 4)
    lab = torch.zeros(2, 1, 4, 4)
    with pytest.raises(_lib.Im2ImError, match="no CPU path"):
        rcps.miss_counts(out, lab, torch.zeros(3))
    with pytest.raises(_lib.Im2ImError, match="no CPU"):
        cm.calibrate_from_outputs(None, out, lab, dict(device="cpu"))
    with pytest.raises(_lib.Im2ImError):
        cm.fraction_missed_loss((lab, lab, lab), lab)
    with pytest.raises(_lib.Im2ImError, match="no CPU path"):
        rcps.head_nested_sets(torch.zeros(2, 2, 1, 4, 4), 1.0, _lib.IM2IM_HEAD_GAUSSIAN)
    with pytest.raises(_lib.Im2ImError, match="no CPU path"):
        rcps.softmax_sets(torch.zeros(2, 50, 1, 4, 4))


def test_product_does_not_import_the_oracle():
    import subprocess
    import sys
    code = ("import sys; import im2im_uq_b200.calibration.calibrate_model, im2im_uq_b200.models.add_uncertainty, core.calibration.calibrate_model;"
            "assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported by product'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "im2im_uq_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|oracle\.|librcps_oracle|rcps_oracle)", text), f
