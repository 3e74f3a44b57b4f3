"""Every IM2IM_* environment switch the product reads is listed in INTEGRATION.md / DESIGN.md / README.md (a switch that
changes which kernel runs and is not written down is a trap for whoever integrates the library)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_environment_switch_is_documented():
    names = set()
    for f in glob.glob(os.path.join(ROOT, "im2im_uq_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "im2im_uq_b200", "csrc", "*.cuh")):
        names |= set(re.findall(r'getenv\("(IM2IM_[A-Z0-9_]+)"\)', open(f).read()))
    for f in glob.glob(os.path.join(ROOT, "im2im_uq_b200", "**", "*.py"), recursive=True):
        names |= set(re.findall(r'environ(?:\.get)?[\(\[]"(IM2IM_[A-Z0-9_]+)"', open(f).read()))
    assert len(names) >= 15
    docs = "".join(open(os.path.join(ROOT, d)).read() for d in ("INTEGRATION.md", "DESIGN.md", "README.md"))
    missing = sorted(n for n in names if n not in docs)
    assert not missing, missing
