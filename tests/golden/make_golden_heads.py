#!/usr/bin/env python
"""Golden fixtures for the non-quantile heads, produced by RUNNING THE UNMODIFIED REFERENCE on CPU.

Run in the authoring container only:  python tests/golden/make_golden_heads.py   (needs /root/reference)

Per head (`uncertainty_type`) the reference's own set function is driven through the reference's own
ModelWithUncertainty / calibrate_model / fraction_missed_loss with an identity trunk, so chosen head outputs reach the
lambda sweep untouched:
  gaussian               core/models/finallayers/gaussian_layer.py:26-34
  residual_magnitude     core/models/finallayers/residual_magnitude_layer.py:28-36
  residual_magnitude_l1  core/models/finallayers/residual_magnitude_l1_layer.py:28-36
  quantiles_l1           core/models/finallayers/quantile_l1_layer.py:34-44
  inn                    core/models/finallayers/inn_layer.py:30-40
  softmax                core/models/finallayers/softmax_layer.py:27-53
and the training losses of the same files (+ core/models/losses/inn.py) give value/gradient known answers.
Files: tests/golden/heads_<name>.npz, tests/golden/head_loss_kats.npz.
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

SET_FNS = {
    "gaussian": lambda: ref.gaussian_layer.gaussian_regression_nested_sets_from_output,
    "residual_magnitude": lambda: ref.residual_magnitude_layer.residual_magnitude_nested_sets_from_output,
    "residual_magnitude_l1": lambda: ref.residual_magnitude_l1_layer.residual_magnitude_l1_nested_sets_from_output,
    "quantiles_l1": lambda: ref.quantile_l1_layer.quantile_regression_l1_nested_sets_from_output,
    "inn": lambda: ref.inn_layer.inn_nested_sets_from_output,
    "softmax": lambda: ref.softmax_layer.softmax_nested_sets_from_output,
}


class _Identity(torch.nn.Module):
    def forward(self, x):
        return x


def _sprinkle(t, g, values, count):
    flat = t.view(-1)
    idx = torch.randperm(flat.numel(), generator=g)[:count * len(values)]
    for i, v in enumerate(values):
        flat[idx[i * count:(i + 1) * count]] = v


def head_outputs(head, seed, n, c, h, w, nasty, noise=1.0, num_softmax=50):
    g = torch.Generator().manual_seed(seed)
    shape = (n, c, h, w)
    pred = torch.rand(shape, generator=g)
    sig = 0.02 + 0.1 * torch.rand(shape, generator=g)
    label = pred + noise * sig * torch.randn(shape, generator=g)
    if head in ("quantiles_l1", "inn"):
        lower = pred - sig * (0.5 + torch.rand(shape, generator=g))
        upper = pred + sig * (0.5 + torch.rand(shape, generator=g))
        if nasty:
            _sprinkle(lower, g, [float("nan"), float("inf"), -float("inf"), 0.7], 3)   # incl. crossed
            _sprinkle(upper, g, [float("nan"), float("inf"), -float("inf"), 0.1], 3)
            _sprinkle(label, g, [float("nan"), float("inf"), 5.0, -5.0], 3)
        out = torch.stack([lower, pred, upper], dim=1)
    elif head in ("residual_magnitude", "residual_magnitude_l1"):
        r = sig * (0.5 + torch.rand(shape, generator=g))
        if nasty:   # the layer's abs() never emits a negative magnitude, but the set function accepts any tensor
            _sprinkle(r, g, [float("nan"), float("inf"), 0.0, -0.0, -0.05, -1e-3, -float("inf")], 4)
            _sprinkle(label, g, [float("nan"), float("inf"), 5.0, -5.0], 3)
            _sprinkle(pred, g, [float("nan"), float("inf"), -float("inf"), 1e6, -1e6], 2)
        out = torch.stack([pred, r], dim=1)
    elif head == "gaussian":
        var = (sig * (0.5 + torch.rand(shape, generator=g))) ** 2
        if nasty:   # relu() never emits a negative variance; sqrt(<0) = nan must propagate like the reference
            _sprinkle(var, g, [float("nan"), float("inf"), 0.0, -0.0, -0.01, 1e-30, 1e30], 4)
            _sprinkle(label, g, [float("nan"), float("inf"), 5.0, -5.0], 3)
            _sprinkle(pred, g, [float("nan"), float("inf"), -float("inf"), 1e6, -1e6], 2)
        out = torch.stack([pred, var], dim=1)
    elif head == "softmax":
        assert c == 1
        classes = torch.linspace(0, 1, num_softmax).view(1, num_softmax, 1, 1, 1)
        centre = (pred + 0.5 * sig * torch.randn(shape, generator=g)).unsqueeze(1)
        width = (0.03 + 1.5 * sig).unsqueeze(1)
        out = -0.5 * ((classes - centre) / width) ** 2 + 0.3 * torch.randn((n, num_softmax, c, h, w), generator=g)
        if nasty:
            _sprinkle(out, g, [float("nan"), 40.0, -40.0], 3)
            _sprinkle(label, g, [float("nan"), 5.0, -5.0, 0.0, 1.0], 3)
            flat_l = label.view(-1)
            flat_l[::11] = (torch.randint(0, num_softmax, (flat_l[::11].numel(),), generator=g).float() / num_softmax)
    else:
        raise ValueError(head)
    return out.contiguous(), label.contiguous()


def torch_softmax_sets(output):
    """Transcription of softmax_layer.py:34-48 with the same torch CPU ops (checked against the reference below)."""
    output = output.softmax(dim=1)
    k = output.shape[1]
    cumsum = torch.cumsum(output, dim=1)
    lq = (cumsum <= 0.05).float().sum(dim=1) / k
    uq = (cumsum <= 0.95).float().sum(dim=1) / k
    pred = torch.argmax(output, dim=1) / k
    lq[pred == lq] -= 1 / k
    uq[pred == uq] += 1 / k
    return torch.stack([lq.clamp(min=0, max=1), pred, uq.clamp(min=0, max=1)], dim=1), cumsum


def run_case(name, head, seed, n, c, h, w, nasty, lam_min, lam_max, num_lambdas, alpha, delta, noise=1.0):
    outputs, labels = head_outputs(head, seed, n, c, h, w, nasty, noise)
    config = dict(alpha=alpha, delta=delta, device="cpu", uncertainty_type=head, minimum_lambda=lam_min,
                  maximum_lambda=lam_max, num_lambdas=num_lambdas, rcps_loss="fraction_missed", dataset="synthetic",
                  batch_size=7, q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0, beta=0.1,
                  num_softmax=50)
    if head == "softmax":  # calibrate_model.py:97-98 reads its own grid keys for this head
        config.update(minimum_lambda_softmax=lam_min, maximum_lambda_softmax=lam_max)
    model = ref.add_uncertainty.ModelWithUncertainty(_Identity(), _Identity(), None, SET_FNS[head](), config)
    dataset = torch.utils.data.TensorDataset(outputs.clone(), labels.clone())
    model, calib_table = ref.calibrate_model.calibrate_model(model, dataset, config)
    lhat = model.lhat.clone()
    lambdas = torch.linspace(lam_min, lam_max, num_lambdas)
    dlambda = lambdas[1] - lambdas[0]
    lam_prime = torch.stack([lam - dlambda for lam in lambdas])
    hit = (lambdas == lhat).nonzero()
    stop_idx = int(hit[0]) if hit.numel() else -1
    loss_fn = ref.calibrate_model.fraction_missed_loss
    dense_prime = torch.zeros(n, num_lambdas)
    dense_grid = torch.zeros(n, num_lambdas)
    with torch.no_grad():
        for j in range(num_lambdas):
            for lo in range(0, n, 4):
                sl = slice(lo, lo + 4) if n - lo != 1 else slice(lo - 1, lo + 1)  # dodge the batch-of-one squeeze quirk
                x, y = outputs[sl], labels[sl]
                a = loss_fn(model.nested_sets_from_output(x.clone(), lam_prime[j]), y)
                b = loss_fn(model.nested_sets_from_output(x.clone(), lambdas[j]), y)
                if n - lo == 1:
                    dense_prime[lo, j], dense_grid[lo, j] = a[-1], b[-1]
                else:
                    dense_prime[sl, j], dense_grid[sl, j] = a, b
        lower, pred, upper = model.nested_sets_from_output(outputs.clone(), lhat)
        lower_mid, _, upper_mid = model.nested_sets_from_output(outputs.clone(), lam_prime[num_lambdas // 2])
    visited = (torch.arange(num_lambdas) >= stop_idx) if stop_idx >= 0 else torch.ones(num_lambdas, dtype=torch.bool)
    assert torch.equal(calib_table[:, visited], dense_prime[:, visited]), name
    assert torch.count_nonzero(calib_table[:, ~visited]) == 0, name
    px = labels[0].numel()
    counts_prime = torch.round(dense_prime.double() * px).to(torch.int32)
    counts_grid = torch.round(dense_grid.double() * px).to(torch.int32)
    assert torch.equal(counts_prime.float() / float(px), dense_prime), name
    assert torch.equal(counts_grid.float() / float(px), dense_grid), name
    extra = {}
    if head == "softmax":
        sets, cumsum = torch_softmax_sets(outputs.clone())
        # the transcription reproduces the reference's edges bit for bit at two lambdas
        for lam, (lo_r, up_r) in ((lhat, (lower, upper)), (lam_prime[num_lambdas // 2], (lower_mid, upper_mid))):
            p = sets[:, 1]
            lo_t = torch.minimum(p - (p - sets[:, 0]).relu() * lam, p - 1e-6)
            up_t = torch.maximum(p + (sets[:, 2] - p).relu() * lam, p + 1e-6)
            assert torch.equal(lo_t.nan_to_num(7.0), lo_r.nan_to_num(7.0)) and torch.equal(up_t.nan_to_num(7.0), up_r.nan_to_num(7.0))
        # distance of every cumulative sum from the two thresholds: pixels closer than a few ulp may legitimately
        # differ between softmax implementations (torch CPU vs torch CUDA vs ours)
        margin = torch.minimum((cumsum - 0.05).abs(), (cumsum - 0.95).abs()).amin(dim=1)
        extra = dict(softmax_sets=sets.numpy(), threshold_margin=margin.numpy())
    np.savez_compressed(
        os.path.join(HERE, f"heads_{name}.npz"), head=head,
        outputs=outputs.numpy(), labels=labels.numpy(), lambdas=lambdas.numpy(), lam_prime=lam_prime.numpy(),
        calib_loss_table=calib_table.numpy(), lhat=lhat.numpy(), stop_idx=np.int64(stop_idx),
        counts_prime=counts_prime.numpy(), counts_grid=counts_grid.numpy(), lower_at_lhat=lower.numpy(),
        upper_at_lhat=upper.numpy(), pred_at_lhat=pred.numpy(), lower_mid=lower_mid.numpy(), upper_mid=upper_mid.numpy(),
        config=json.dumps(dict(alpha=alpha, delta=delta, minimum_lambda=lam_min, maximum_lambda=lam_max,
                               num_lambdas=num_lambdas, seed=seed, nasty=nasty, noise=noise, uncertainty_type=head,
                               beta=0.1, num_softmax=50)), **extra)
    print(f"[golden] {name} ({head}): N={n} C={c} {h}x{w} L={num_lambdas} stop_idx={stop_idx} lhat={float(lhat):.9g} "
          f"visited={int(visited.sum())}")


def loss_kats():
    """Value + gradient of every head's training loss (reference autograd on CPU)."""
    g = torch.Generator().manual_seed(21)
    fns = dict(gaussian=ref.gaussian_layer.gaussian_regression_loss_fn,
               residual_magnitude=ref.residual_magnitude_layer.residual_magnitude_loss_fn,
               residual_magnitude_l1=ref.residual_magnitude_l1_layer.residual_magnitude_l1_loss_fn,
               quantiles_l1=ref.quantile_l1_layer.quantile_regression_l1_loss_fn,
               inn=ref.inn_layer.inn_loss_fn, softmax=ref.softmax_layer.softmax_loss_fn)
    cases = {}
    for head, fn in fns.items():
        for tag, (b, c, h, w) in dict(a=(4, 1, 16, 16), b=(3, 1, 9, 5)).items():
            params = dict(q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0, beta=0.1,
                          num_softmax=50, device="cpu")
            if tag == "b":
                params.update(q_lo=0.1, q_hi=0.8, q_lo_weight=0.5, q_hi_weight=2.0, mse_weight=0.25, beta=0.3)
            target = torch.rand(b, c, h, w, generator=g)
            if head == "softmax":
                pred = torch.randn(b, 50, c, h, w, generator=g)
                target.view(-1)[::5] = torch.linspace(0, 1, 50)[torch.randint(0, 50, (target.view(-1)[::5].numel(),), generator=g)]
                target.view(-1)[1::17] = 1.5   # beyond the last class edge -> clamped to the last class (:22)
            else:
                planes = 3 if head in ("quantiles_l1", "inn") else 2
                pred = torch.randn(b, planes, c, h, w, generator=g)
                if head == "gaussian":
                    pred[:, 1] = pred[:, 1].abs() + 0.05       # the layer's relu output; keep away from the eps clamp
                    pred[:, 1, :, ::3, ::3] = 0.0               # relu zeros -> clamped to eps=1e-6 by GaussianNLLLoss
                if head.startswith("residual"):
                    pred[:, 1] = pred[:, 1].abs()
            pred_g = pred.clone().requires_grad_(True)
            loss = fn(pred_g, target.clone(), params)
            loss.backward()
            key = f"{head}_{tag}"
            cases[key + "_pred"] = pred.numpy(); cases[key + "_target"] = target.numpy()
            cases[key + "_loss"] = loss.detach().numpy(); cases[key + "_grad"] = pred_g.grad.numpy()
            cases[key + "_params"] = json.dumps(params)
    np.savez_compressed(os.path.join(HERE, "head_loss_kats.npz"), **cases)
    print(f"[golden] head losses: {len(fns) * 2} cases")


def loss_table_trials():
    """evaluate_from_loss_table (calibrate_model.py:62-74) as plot_risks drives it (experiments/*/plot.py:133-136):
    seeded trials on saved dense tables; pins the RNG consumption and the `<= delta` column choice."""
    import warnings
    cases = {}
    for name, n, delta, trials in (("fastmri_small", 24, 0.1, 6), ("temca_small", 40, 0.3, 6), ("bsbcm_grid", 12, 0.2, 3)):
        table = torch.from_numpy(np.load(os.path.join(HERE, f"rcps_{name}.npz"))["dense_grid"])
        torch.manual_seed(1234)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            risks = torch.stack([ref.calibrate_model.evaluate_from_loss_table(table, n, 0.1, delta) for _ in range(trials)])
        cases[name + "_risks"] = risks.numpy()
        cases[name + "_args"] = json.dumps(dict(n=n, alpha=0.1, delta=delta, trials=trials, seed=1234))
    np.savez_compressed(os.path.join(HERE, "loss_table_trials.npz"), **cases)
    print("[golden] loss-table trials:", {k: v.tolist() for k, v in cases.items() if k.endswith("_risks")})


if __name__ == "__main__":
    only = sys.argv[1:]
    if only == ["trials"]:
        loss_table_trials()
        sys.exit(0)
    if not only:
        loss_kats()
        loss_table_trials()
    _run_case = run_case
    run_case = lambda name, *a, **k: _run_case(name, *a, **k) if (not only or name in only) else None  # noqa: E731
    #        name               head                     seed n   c  h   w   nasty  lmin lmax L     alpha delta
    run_case("gaussian_small",  "gaussian",              10,  40, 1, 24, 24, False, 0.0, 6.0, 400,  0.1,  0.1)
    run_case("gaussian_nasty",  "gaussian",              11,  33, 2, 11, 7,  True, -1.0, 5.0, 120,  0.3,  0.1)
    run_case("residual_small",  "residual_magnitude",    12,  40, 1, 24, 24, False, 0.0, 6.0, 400,  0.1,  0.1)
    run_case("residual_nasty",  "residual_magnitude",    13,  33, 3, 9,  7,  True, -2.0, 4.0, 150,  0.4,  0.1)
    run_case("residual_l1",     "residual_magnitude_l1", 14,  21, 1, 16, 12, True,  0.0, 6.0, 64,   0.2,  0.1)
    run_case("quantiles_l1",    "quantiles_l1",          15,  26, 1, 16, 16, True,  0.0, 6.0, 100,  0.2,  0.1)
    run_case("inn",             "inn",                   16,  26, 2, 8,  10, True,  0.0, 6.0, 100,  0.2,  0.1)
    run_case("softmax_small",   "softmax",               17,  30, 1, 16, 16, False, 0.0, 8.0, 200,  0.2,  0.1)
    run_case("softmax_nasty",   "softmax",               18,  22, 1, 12, 10, True,  0.0, 8.0, 80,   0.3,  0.1)
