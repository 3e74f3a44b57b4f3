#!/usr/bin/env python
"""Golden fixtures for SURVEY.md §8(f1): the reference's OWN ``get_rcps_metrics_from_outputs``
(core/calibration/calibrate_model.py:31-60) run unmodified on CPU over the head outputs of existing rcps_* fixtures.

Run in the authoring container only:  python tests/golden/make_golden_metrics.py   (needs /root/reference)

The function draws ``np.random.choice`` once per batch of 64 (:44) and one ``torch.rand`` (:51); both generators are
seeded here (np.random.seed(seed); torch.manual_seed(seed)) and the seed is stored, so a replacement that makes the same
calls in the same order must reproduce every returned value: losses, sizes, spearman, stratified_risks, mse and the
(H, W) spatial miscoverage map.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_import import import_reference  # noqa: E402

ref = import_reference()
torch.set_num_threads(8)


class _Identity(torch.nn.Module):
    def forward(self, x):
        return x


def run(name, source, seed, lhat=None):
    g = np.load(os.path.join(HERE, f"rcps_{source}.npz"))
    outputs, labels = torch.from_numpy(g["outputs"]).clone(), torch.from_numpy(g["labels"]).clone()
    params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
    model = ref.add_uncertainty.ModelWithUncertainty(
        _Identity(), _Identity(), ref.quantile_layer.quantile_regression_loss_fn,
        ref.quantile_layer.quantile_regression_nested_sets_from_output, params)
    lh = torch.tensor(g["lhat"]) if lhat is None else torch.tensor(lhat, dtype=torch.float32)
    model.set_lhat(lh)
    ds = torch.utils.data.TensorDataset(outputs, labels)
    np.random.seed(seed)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
        losses, sizes, spearman, strat, mse, spatial = ref.calibrate_model.get_rcps_metrics_from_outputs(
            model, ds, ref.calibrate_model.fraction_missed_loss, "cpu")
    np.savez_compressed(os.path.join(HERE, f"metrics_{name}.npz"), source=source, seed=np.int64(seed),
                        lhat=lh.numpy(), losses=losses.numpy(), sizes=sizes.numpy(), spearman=np.float64(spearman),
                        stratified_risks=strat.numpy(), mse=np.float64(mse), spatial_miscoverage=np.asarray(spatial))
    print(f"[golden] metrics_{name}: N={outputs.shape[0]} lhat={float(lh):.6g} risk={float(losses.mean()):.6g} "
          f"spearman={float(spearman):.6g} mse={mse:.6g} spatial {np.asarray(spatial).shape}")


if __name__ == "__main__":
    run("fastmri_small", "fastmri_small", 11)          # one batch of 48
    run("batch65", "batch65", 12)                      # 66 images: batches of 64 + 2 (two np.random.choice calls)
    run("temca_small", "temca_small", 13)              # 70 images, non-square 16x24
    run("bsbcm_grid", "bsbcm_grid", 14)                # 2 channels: spatial map averaged over the channel axis
    run("fastmri_lam1", "fastmri_small", 15, lhat=1.0)  # a lambda with a large risk
