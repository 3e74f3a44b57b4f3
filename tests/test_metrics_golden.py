"""``get_rcps_metrics_from_outputs`` against fixtures produced by the reference's OWN function
(core/calibration/calibrate_model.py:31-60, run unmodified by tests/golden/make_golden_metrics.py with seeded RNGs).

CPU half: the oracle's numpy restatement reproduces every returned value.  GPU half (-m gpu): the product function,
through the C ABI kernels, reproduces them too - losses, sampled sizes, stratified risks and the spatial miscoverage
map bit for bit; Spearman and mse to 1e-12 relative (float64 host arithmetic on identical fp32 inputs)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden

CASES = sorted(os.path.basename(f)[len("metrics_"):-len(".npz")] for f in glob.glob(os.path.join(GOLDEN, "metrics_*.npz")))


def _load(name):
    g = np.load(os.path.join(GOLDEN, f"metrics_{name}.npz"))
    src = load_golden(str(g["source"]))
    return g, src


def _check(g, got, exact_float=True):
    losses, sizes, spearman, strat, mse, spatial = got
    assert np.array_equal(np.asarray(losses.cpu()), g["losses"])
    assert np.array_equal(np.asarray(sizes.cpu()), g["sizes"])
    assert np.array_equal(np.asarray(strat), g["stratified_risks"], equal_nan=True)
    assert np.array_equal(np.asarray(spatial), g["spatial_miscoverage"])
    assert np.asarray(spatial).dtype == g["spatial_miscoverage"].dtype
    assert abs(float(spearman) - float(g["spearman"])) <= 1e-12 * max(1.0, abs(float(g["spearman"])))
    assert abs(float(mse) - float(g["mse"])) <= 1e-12 * abs(float(g["mse"]))


def test_fixtures_exist():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_oracle_metrics_match_reference(name):
    from oracle import rcps_oracle as orc
    g, src = _load(name)
    got = orc.np_metrics_at_lambda(src["outputs"], src["labels"], g["lhat"], seed=int(g["seed"]))
    _check(g, got)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_native_metrics_match_reference(name):
    from core.calibration import calibrate_model as cm
    from core.models.add_uncertainty import ModelWithUncertainty
    from core.models.finallayers.quantile_layer import (quantile_regression_loss_fn,
                                                        quantile_regression_nested_sets_from_output)
    from im2im_uq_b200 import _lib
    g, src = _load(name)

    class _Id(torch.nn.Module):
        def forward(self, x):
            return x

    model = ModelWithUncertainty(_Id(), _Id(), quantile_regression_loss_fn,
                                 quantile_regression_nested_sets_from_output, src["config"])
    model.set_lhat(torch.tensor(g["lhat"]))
    ds = torch.utils.data.TensorDataset(torch.from_numpy(src["outputs"]).clone(), torch.from_numpy(src["labels"]).clone())
    np.random.seed(int(g["seed"]))
    torch.manual_seed(int(g["seed"]))
    before = _lib.launch_count()
    got = cm.get_rcps_metrics_from_outputs(model, ds, cm.fraction_missed_loss, "cuda:0")
    assert _lib.launch_count() - before >= 3          # miss counts, endpoints at the sampled pixels, miss map
    _check(g, got)
