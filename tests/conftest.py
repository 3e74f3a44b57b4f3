import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    return sorted(os.path.basename(f)[len("rcps_"):-len(".npz")] for f in glob.glob(os.path.join(GOLDEN, "rcps_*.npz")))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, f"rcps_{name}.npz"))
    d = {k: g[k] for k in g.files}
    d["config"] = json.loads(str(d["config"]))
    d["config"].update(uncertainty_type="quantiles", rcps_loss="fraction_missed", dataset="synthetic",
                       batch_size=7, q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
    return d


@pytest.fixture(params=golden_cases())
def golden(request):
    d = load_golden(request.param)
    d["name"] = request.param
    return d


def synth_scores(seed, n, c, h, w, device="cpu", noise=1.0):
    """SURVEY.md §8c probe recipe: lambda-hat lands mid-grid on the fastmri grid."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    shape = (n, c, h, w)
    pred = torch.rand(shape, generator=g, device=device)
    sig = 0.02 + 0.1 * torch.rand(shape, generator=g, device=device)
    lower = pred - sig * (0.5 + torch.rand(shape, generator=g, device=device))
    upper = pred + sig * (0.5 + torch.rand(shape, generator=g, device=device))
    label = pred + noise * sig * torch.randn(shape, generator=g, device=device)
    return torch.stack([lower, pred, upper], dim=1).contiguous(), label.contiguous()
