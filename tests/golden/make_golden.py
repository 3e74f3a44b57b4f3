#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE on CPU.

Run in the authoring container only:  python tests/golden/make_golden.py
(needs /root/reference; the GPU box never runs this - it only reads the committed .npz/.json).

What is pinned (reference file:line):
  * HB_mu_plus known answers                      core/calibration/bounds.py:17-29
  * calibrate_model end-to-end: loss table, lhat   core/calibration/calibrate_model.py:89-145
    driven through an identity "model" so that chosen head outputs reach the lambda sweep untouched
  * dense per-lambda loss tables (get_loss_table inner loop, core/scripts/eval.py:115-125) computed with the
    reference's own nested_sets_from_output (add_uncertainty.py:33-38) + fraction_missed_loss
    (calibrate_model.py:76-80)
  * interval endpoints at lhat                     add_uncertainty.py:33-38, quantile_layer.py:34-44
  * quantile training loss                         quantile_layer.py:23-32, losses/pinball.py:12-24
"""
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


def head_outputs(seed, n, c, h, w, kind="probe", noise=1.0):
    """Synthetic quantile-head outputs + labels (SURVEY.md §8c probe recipe and nastier variants)."""
    g = torch.Generator().manual_seed(seed)
    shape = (n, c, h, w)
    pred = torch.rand(shape, generator=g)
    sig = 0.02 + 0.1 * torch.rand(shape, generator=g)
    lower = pred - sig * (0.5 + torch.rand(shape, generator=g))
    upper = pred + sig * (0.5 + torch.rand(shape, generator=g))
    label = pred + noise * sig * torch.randn(shape, generator=g)
    if kind == "nasty":
        flat = lambda t: t.view(-1)
        m = flat(pred).numel()
        idx = torch.randperm(m, generator=g)
        k = max(m // 16, 1)
        # crossed quantiles, zero width, exact ties label==pred / label==edge, huge and tiny magnitudes, nan/inf
        flat(lower)[idx[0 * k:1 * k]] = flat(pred)[idx[0 * k:1 * k]] + 0.05          # lower above pred
        flat(upper)[idx[1 * k:2 * k]] = flat(pred)[idx[1 * k:2 * k]] - 0.05          # upper below pred
        flat(lower)[idx[2 * k:3 * k]] = flat(pred)[idx[2 * k:3 * k]]                 # zero width low
        flat(upper)[idx[2 * k:3 * k]] = flat(pred)[idx[2 * k:3 * k]]                 # zero width high
        flat(label)[idx[3 * k:4 * k]] = flat(pred)[idx[3 * k:4 * k]]                 # label == pred
        flat(label)[idx[4 * k:5 * k]] = flat(upper)[idx[4 * k:5 * k]]                # label == upper
        flat(label)[idx[5 * k:6 * k]] = flat(lower)[idx[5 * k:6 * k]]                # label == lower
        sl = idx[6 * k:7 * k]
        flat(pred)[sl] *= 1e6; flat(lower)[sl] *= 1e6; flat(upper)[sl] *= 1e6; flat(label)[sl] *= 1e6
        sl = idx[7 * k:8 * k]
        flat(pred)[sl] *= 1e-6; flat(lower)[sl] *= 1e-6; flat(upper)[sl] *= 1e-6; flat(label)[sl] *= 1e-6
        sp = idx[8 * k:8 * k + 12]
        vals = [float("nan"), float("inf"), -float("inf")]
        for t_i, t in enumerate((lower, pred, upper, label)):
            for v_i, v in enumerate(vals):
                flat(t)[sp[t_i * 3 + v_i]] = v
        flat(label)[idx[9 * k:10 * k]] += 5.0                                         # far above: missed for most lambdas
        flat(label)[idx[10 * k:11 * k]] -= 5.0                                        # far below
        flat(pred)[idx[11 * k:12 * k]] = -flat(pred)[idx[11 * k:12 * k]]              # negative preds w/ inconsistent edges
    elif kind == "covered":
        # every label strictly inside the heuristic interval scaled by the top lambda -> top-of-grid risk exactly 0
        label = pred + 0.25 * sig * (2 * torch.rand(shape, generator=g) - 1)
    outputs = torch.stack([lower, pred, upper], dim=1).contiguous()  # (N,3,C,H,W)
    return outputs, label.contiguous()


class _Identity(torch.nn.Module):
    def forward(self, x):
        return x


def make_model(params):
    """Reference ModelWithUncertainty whose forward is the identity, so dataset inputs ARE the head outputs."""
    return ref.add_uncertainty.ModelWithUncertainty(
        _Identity(), _Identity(), ref.quantile_layer.quantile_regression_loss_fn,
        ref.quantile_layer.quantile_regression_nested_sets_from_output, params)


def run_case(name, seed, n, c, h, w, kind, lam_min, lam_max, num_lambdas, alpha, delta, noise=1.0):
    outputs, labels = head_outputs(seed, n, c, h, w, kind, noise)
    config = dict(alpha=alpha, delta=delta, device="cpu", uncertainty_type="quantiles", minimum_lambda=lam_min,
                  maximum_lambda=lam_max, num_lambdas=num_lambdas, rcps_loss="fraction_missed", dataset="synthetic",
                  batch_size=7, q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
    model = make_model(config)
    dataset = torch.utils.data.TensorDataset(outputs.clone(), labels.clone())
    model, calib_table = ref.calibrate_model.calibrate_model(model, dataset, config)
    lhat = model.lhat.clone()
    lambdas = torch.linspace(lam_min, lam_max, num_lambdas)
    dlambda = lambdas[1] - lambdas[0]
    lam_prime = torch.stack([lam - dlambda for lam in lambdas])  # what calibrate_model.py:135 evaluates
    # stop index: the column of lhat; -1 if the sweep never stopped (lhat keeps its default, calibrate_model.py:131)
    hit = (lambdas == lhat).nonzero()
    stop_idx = int(hit[0]) if hit.numel() else -1
    # dense tables (no early stop) at lam' (calibration grid) and at lambdas (eval.py:122-124 grid)
    loss_fn = ref.calibrate_model.fraction_missed_loss
    dense_prime = torch.zeros(n, num_lambdas)
    dense_grid = torch.zeros(n, num_lambdas)
    with torch.no_grad():
        for j in range(num_lambdas):
            for lo in range(0, n, 4):
                x = outputs[lo:lo + 4].clone(); y = labels[lo:lo + 4]
                if x.shape[0] == 1:  # reference's .squeeze() drops a batch of one - avoid it
                    x = outputs[lo - 1:lo + 1].clone(); y = labels[lo - 1:lo + 1]
                    dense_prime[lo, j] = loss_fn(model.nested_sets_from_output(x.clone(), lam_prime[j]), y)[-1]
                    dense_grid[lo, j] = loss_fn(model.nested_sets_from_output(x.clone(), lambdas[j]), y)[-1]
                    continue
                dense_prime[lo:lo + 4, j] = loss_fn(model.nested_sets_from_output(x.clone(), lam_prime[j]), y)
                dense_grid[lo:lo + 4, j] = loss_fn(model.nested_sets_from_output(x.clone(), lambdas[j]), y)
        lower, pred, upper = model.nested_sets_from_output(outputs.clone(), lhat)
    visited = (torch.arange(num_lambdas) >= stop_idx) if stop_idx >= 0 else torch.ones(num_lambdas, dtype=torch.bool)
    # self-consistency of what we are about to freeze
    assert torch.equal(calib_table[:, visited], dense_prime[:, visited]), name
    assert torch.count_nonzero(calib_table[:, ~visited]) == 0, name
    px = c * h * w
    counts_prime = torch.round(dense_prime.double() * px).to(torch.int32)
    counts_grid = torch.round(dense_grid.double() * px).to(torch.int32)
    assert torch.equal((counts_prime.float() / float(px)), dense_prime), name
    assert torch.equal((counts_grid.float() / float(px)), dense_grid), name
    np.savez_compressed(
        os.path.join(HERE, f"rcps_{name}.npz"),
        outputs=outputs.numpy(), labels=labels.numpy(), lambdas=lambdas.numpy(), lam_prime=lam_prime.numpy(),
        calib_loss_table=calib_table.numpy(), lhat=lhat.numpy(), stop_idx=np.int64(stop_idx),
        dense_prime=dense_prime.numpy(), dense_grid=dense_grid.numpy(), counts_prime=counts_prime.numpy(),
        counts_grid=counts_grid.numpy(), lower_at_lhat=lower.numpy(), upper_at_lhat=upper.numpy(),
        pred_at_lhat=pred.numpy(),
        config=json.dumps(dict(alpha=alpha, delta=delta, minimum_lambda=lam_min, maximum_lambda=lam_max,
                               num_lambdas=num_lambdas, seed=seed, kind=kind, noise=noise)))
    print(f"[golden] {name}: N={n} C={c} {h}x{w} L={num_lambdas} stop_idx={stop_idx} lhat={float(lhat):.9g} "
          f"visited={int(visited.sum())}")


def hb_kats():
    rows = []
    for n in (1, 7, 32, 100, 1000, 4000, 10000, 50000):
        for delta in (0.1, 0.05, 0.001):
            for muhat in (0.0, 1e-6, 1e-3, 0.01, 0.05, 0.0999, 0.1, 0.25, 0.5, 0.9, 0.999999, 1.0):
                rows.append(dict(muhat=muhat, n=n, delta=delta, value=float(ref.bounds.HB_mu_plus(muhat, n, delta))))
    # fp32-valued muhat exactly as calibrate_model passes them (Rhat.item())
    g = torch.Generator().manual_seed(3)
    for v in torch.rand(40, generator=g).mul(0.3).tolist():
        rows.append(dict(muhat=v, n=1000, delta=0.1, value=float(ref.bounds.HB_mu_plus(v, 1000, 0.1))))
    with open(os.path.join(HERE, "hb_mu_plus_kats.json"), "w") as f:
        json.dump(rows, f, indent=0)
    print(f"[golden] HB_mu_plus: {len(rows)} known answers")


def quantile_loss_kats():
    g = torch.Generator().manual_seed(11)
    cases = {}
    for name, (b, c, h, w) in dict(a=(4, 1, 16, 16), b=(3, 2, 9, 5), c=(2, 1, 32, 32)).items():
        pred = torch.randn(b, 3, c, h, w, generator=g)
        target = torch.randn(b, c, h, w, generator=g)
        target.view(-1)[::7] = pred[:, 0].reshape(-1)[::7]  # exact ties -> zero pinball contribution
        params = dict(q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
        if name == "b":
            params = dict(q_lo=0.1, q_hi=0.8, q_lo_weight=0.5, q_hi_weight=2.0, mse_weight=0.25)
        pred_g = pred.clone().requires_grad_(True)
        loss = ref.quantile_layer.quantile_regression_loss_fn(pred_g, target, params)
        loss.backward()
        cases[f"{name}_pred"] = pred.numpy(); cases[f"{name}_target"] = target.numpy()
        cases[f"{name}_loss"] = loss.detach().numpy(); cases[f"{name}_grad"] = pred_g.grad.numpy()
        cases[f"{name}_params"] = json.dumps(params)
    np.savez_compressed(os.path.join(HERE, "quantile_loss_kats.npz"), **cases)
    print("[golden] quantile loss: 3 cases")


def unet_forward_kat():
    """Reference UNet(1,1)+quantile head: seeded init, BN running stats moved by 3 train-mode batches, eval forward.
    core/models/trunks/unet.py:10-46, core/models/finallayers/quantile_layer.py:8-21, add_uncertainty.py:25-27."""
    params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
    torch.manual_seed(0)
    model = ref.add_uncertainty.add_uncertainty(ref.unet.UNet(1, 1), params)
    g = torch.Generator().manual_seed(1)
    model.train()
    with torch.no_grad():
        for _ in range(3):
            model(torch.randn(4, 1, 32, 32, generator=g))
    model.eval()
    x = torch.randn(2, 1, 48, 32, generator=g)
    with torch.no_grad():
        y = model(x)
    sd = model.state_dict()
    checksum = float(sum(v.double().abs().sum() for k, v in sd.items() if v is not None and v.dtype.is_floating_point))
    np.savez_compressed(os.path.join(HERE, "unet_forward_kat.npz"), x=x.numpy(), y=y.numpy(),
                        state_checksum=np.float64(checksum), n_params=np.int64(sum(p.numel() for p in model.parameters())),
                        keys=json.dumps(list(sd.keys())))
    print(f"[golden] unet forward: y {tuple(y.shape)} checksum {checksum:.6f}")


if __name__ == "__main__":
    unet_forward_kat()
    hb_kats()
    quantile_loss_kats()
    #        name            seed  n   c  h   w   kind      lam_min lam_max L     alpha delta
    run_case("fastmri_small", 0,   48, 1, 32, 32, "probe",   0.0,   6.0,   1000, 0.1,  0.1)
    run_case("temca_small",   1,   70, 1, 16, 24, "probe",   7.0,   10.0,  100,  0.3,  0.1, noise=6.0)
    run_case("nasty_ragged",  2,   37, 3, 13, 7,  "nasty",   0.0,   6.0,   250,  0.5,  0.1)
    run_case("top_risk_zero", 3,   20, 1, 16, 16, "covered", 0.0,   6.0,   64,   0.1,  0.1)
    run_case("never_stops",   4,   16, 1, 8,  8,  "nasty",   0.0,   6.0,   40,   0.9999, 0.1)
    run_case("bsbcm_grid",    5,   24, 2, 10, 10, "probe",   0.0,   6.0,   2000, 0.25, 0.1)
    run_case("neg_grid",      6,   30, 1, 12, 12, "nasty",  -1.0,   2.0,   33,   0.6,  0.2)
    run_case("batch65",       7,   66, 1, 8,  8,  "probe",   0.0,   4.0,   50,   0.2,  0.05)
