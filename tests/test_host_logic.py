"""CPU: host-side mirror of the reference interface - bounds, lambda grid, stopping-rule replay, loss-table eval.

Counts come from the oracle here (the CUDA kernel is exercised by the -m gpu tests); what is under test is the host
logic that turns one-pass integer counts into the reference's lhat / table.
"""
import contextlib
import io
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, synth_scores
from im2im_uq_b200.calibration import bounds, sweep
from oracle import rcps_oracle as orc


def test_bounds_known_answers():
    kats = json.load(open(os.path.join(GOLDEN, "hb_mu_plus_kats.json")))
    with contextlib.redirect_stdout(io.StringIO()):
        for k in kats:
            assert bounds.HB_mu_plus(k["muhat"], k["n"], k["delta"]) == k["value"], k


def test_bounds_zero_risk_quirk_and_helpers(capsys):
    assert bounds.HB_mu_plus(0.0, 1000, 0.1) == 1.0  # reference's exception path
    assert "BRENTQ RUNTIME ERROR" in capsys.readouterr().out
    assert bounds.HB_mu_plus(1.0, 100, 0.1) == 1
    assert bounds.h1(0.3, 0.3) == 0.0
    assert bounds.hoeffding_plus(0.5, 0.2, 10) < 0 and bounds.hoeffding_plus(0.2, 0.5, 10) == 0.0


@pytest.mark.parametrize("n,alpha,delta", [(48, 0.1, 0.1), (1000, 0.1, 0.1), (10000, 0.1, 0.1), (70, 0.3, 0.1),
                                           (16, 0.1, 0.1), (4000, 0.05, 0.001)])
def test_hb_bracket_is_the_level_set(n, alpha, delta):
    r_lo, r_hi = bounds.hb_stop_bracket(n, alpha, delta)
    if r_lo == 0.0:
        assert bounds._hb_quiet(1e-9, n, delta) > alpha
        return
    assert r_hi - r_lo < 1e-9
    assert bounds._hb_quiet(r_lo, n, delta) <= alpha < bounds._hb_quiet(r_hi, n, delta)
    rng = np.random.default_rng(0)
    for m in rng.uniform(1e-6, 0.6, 40):  # monotone: one crossing
        assert (bounds._hb_quiet(m, n, delta) > alpha) == (m >= r_hi) or r_lo < m < r_hi


def test_lambda_grid_matches_reference(golden):
    lambdas, dlambda, lam_prime, default_lhat = sweep.lambda_grid(golden["config"])
    assert np.array_equal(lambdas.numpy(), golden["lambdas"])
    assert np.array_equal(lam_prime.numpy(), golden["lam_prime"])  # elementwise == the per-step `lam - dlambda`
    if int(golden["stop_idx"]) < 0:
        assert np.float32(default_lhat.numpy()) == golden["lhat"]


def _sweep_with_oracle_counts(g, **kw):
    cfg = g["config"]
    counts = torch.from_numpy(orc.c_miss_table(g["outputs"], g["labels"], g["lam_prime"]))
    totals = counts.sum(0, dtype=torch.int64)
    px = int(np.prod(g["outputs"].shape[2:]))
    stats = {}
    lhat, stop, visited = sweep.sweep_from_counts(counts, totals, px, cfg, lambda col: col.float() / float(px),
                                                  stats=stats, **kw)
    table = (counts.float() / float(px)) * visited[None, :].float()
    return lhat, stop, visited, table, stats


def test_stop_rule_matches_reference(golden):
    lhat, stop, visited, table, stats = _sweep_with_oracle_counts(golden)
    assert stop == int(golden["stop_idx"])
    assert np.float32(lhat.numpy()) == golden["lhat"]
    assert np.array_equal(table.numpy(), golden["calib_loss_table"])
    assert stats["screened"]
    assert stats["replayed_columns"] <= 3, stats  # the point of the screening: almost no exact replays


def test_screened_scan_equals_exhaustive_replay(golden):
    a = _sweep_with_oracle_counts(golden)
    b = _sweep_with_oracle_counts(golden, ascending=False)  # replays every column like the reference's loop
    assert a[1] == b[1] and torch.equal(a[0], b[0])
    assert not b[4]["screened"]


@pytest.mark.parametrize("seed", range(6))
def test_screening_never_disagrees_on_random_sets(seed):
    """Property: for random calibration sets the screened decision == literal linear scan (reference semantics)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(20, 400))
    alpha = float(rng.choice([0.05, 0.1, 0.2, 0.3]))
    out, lab = synth_scores(seed, n, 1, 8, 8, noise=float(rng.choice([1.0, 2.0])))
    cfg = dict(uncertainty_type="quantiles", minimum_lambda=0.0, maximum_lambda=6.0, num_lambdas=200, alpha=alpha,
               delta=0.1)
    lambdas, dl, lam_prime, _ = sweep.lambda_grid(cfg)
    counts = torch.from_numpy(orc.c_miss_table(out.numpy(), lab.numpy(), lam_prime.numpy()))
    px = 64
    got = sweep.sweep_from_counts(counts, counts.sum(0, dtype=torch.int64), px, cfg, lambda c: c.float() / float(px))
    ref_lhat, ref_stop, _ = orc.calibrate_sweep(out.numpy(), lab.numpy(), 0.0, 6.0, 200, alpha, 0.1)
    assert got[1] == ref_stop and torch.equal(got[0], ref_lhat)


def test_visited_mask_with_duplicate_lambdas():
    lambdas = torch.tensor([0.0, 1.0, 1.0, 2.0])
    assert sweep.visited_mask(lambdas, 2).tolist() == [False, True, True, True]
    assert sweep.visited_mask(lambdas, -1).tolist() == [True] * 4
    assert sweep.visited_mask(lambdas, 3).tolist() == [False, False, False, True]


def test_one_point_grid_raises_like_reference():
    with pytest.raises(IndexError):
        sweep.lambda_grid(dict(uncertainty_type="quantiles", minimum_lambda=0.0, maximum_lambda=1.0, num_lambdas=1))


def test_evaluate_from_loss_table_matches_literal_restatement():
    from im2im_uq_b200.calibration.calibrate_model import evaluate_from_loss_table
    g = np.load(os.path.join(GOLDEN, "rcps_fastmri_small.npz"))
    table = torch.from_numpy(g["dense_grid"])
    for n, alpha, delta, seed in [(24, 0.1, 0.1, 0), (30, 0.1, 0.3, 1), (10, 0.1, 0.05, 2)]:
        torch.manual_seed(seed)
        got = evaluate_from_loss_table(table, n, alpha, delta)
        # literal restatement of calibrate_model.py:62-74 with the oracle's bound
        torch.manual_seed(seed)
        perm = torch.randperm(table.shape[0])
        t = table[perm]
        calib, val = t[:n], t[n:]
        rhats = calib.mean(dim=0)
        plus = torch.tensor([orc.hb_mu_plus(r, n, delta) for r in rhats])
        nz = (plus <= delta).nonzero()
        idx = nz[0] if nz.numel() else 0
        assert torch.equal(got, val[:, idx].mean()), (n, delta)


def test_get_rcps_loss_fn_contract():
    from im2im_uq_b200.calibration.calibrate_model import fraction_missed_loss, get_rcps_loss_fn
    assert get_rcps_loss_fn({"rcps_loss": "fraction_missed"}) is fraction_missed_loss
    with pytest.raises(NotImplementedError):
        get_rcps_loss_fn({"rcps_loss": "something_else"})


def test_model_surface_and_errors():
    from core.models.add_uncertainty import add_uncertainty
    from core.models.trunks.unet import UNet
    params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
    m = add_uncertainty(UNet(1, 1), params)
    assert m.lhat is None and "lhat" in dict(m.named_buffers(recurse=False)) or m.lhat is None
    assert sum(p.numel() for p in m.parameters()) == 17269123
    with pytest.raises(Exception, match="You have to specify lambda"):
        m.nested_sets((torch.zeros(1, 1, 16, 16),))
    m.set_lhat(torch.tensor(1.5))
    assert float(m.lhat) == 1.5 and "lhat" in m.state_dict()
    with pytest.raises(NotImplementedError):
        add_uncertainty(UNet(1, 1), dict(params, uncertainty_type="no_such_head"))
    y = m(torch.zeros(2, 1, 16, 16))
    assert tuple(y.shape) == (2, 3, 1, 16, 16)


def test_quantile_loss_matches_reference_kats():
    from core.models.finallayers.quantile_layer import quantile_regression_loss_fn
    g = np.load(os.path.join(GOLDEN, "quantile_loss_kats.npz"))
    for name in "abc":
        pred = torch.from_numpy(g[f"{name}_pred"]).requires_grad_(True)
        target = torch.from_numpy(g[f"{name}_target"])
        params = json.loads(str(g[f"{name}_params"]))
        loss = quantile_regression_loss_fn(pred, target, params)
        loss.backward()
        np.testing.assert_allclose(loss.detach().numpy(), g[f"{name}_loss"], rtol=1e-6)
        np.testing.assert_allclose(pred.grad.numpy(), g[f"{name}_grad"], rtol=1e-6, atol=1e-9)


def test_loss_table_trials_match_reference_and_wire_format(tmp_path):
    """plot_risks' trial loop over evaluate_from_loss_table against values produced by the unmodified reference
    (tests/golden/make_golden_heads.py::loss_table_trials), and the router's .pth wire format round trip."""
    import json
    import warnings
    from conftest import load_golden
    from im2im_uq_b200.scripts import eval as ev
    kats = np.load(os.path.join(GOLDEN, "loss_table_trials.npz"))
    for name in ("fastmri_small", "temca_small", "bsbcm_grid"):
        args = json.loads(str(kats[name + "_args"]))
        table = torch.from_numpy(load_golden(name)["dense_grid"])
        torch.manual_seed(args["seed"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            risks = ev.evaluate_loss_table_trials(table, args["n"], args["alpha"], args["delta"], args["trials"])
        assert np.array_equal(risks.numpy(), kats[name + "_risks"]), name
    g = load_golden("batch65")
    calib, val = torch.from_numpy(g["calib_loss_table"]), torch.from_numpy(g["dense_grid"])
    path = str(tmp_path / ev.loss_table_filename(dict(dataset="fastmri", uncertainty_type="quantiles", batch_size=78,
                                                       lr=0.0001, input_normalization="standard",
                                                       output_normalization="min-max")))
    assert path.endswith("loss_table_fastmri_quantiles_78_0.0001_standard_min-max.pth")
    ev.save_loss_tables(calib, val, path)
    back = ev.load_loss_table(path)
    assert back.shape == (calib.shape[0] + val.shape[0], calib.shape[1])
    assert torch.equal(back[:calib.shape[0]], calib) and torch.equal(back[calib.shape[0]:], val)
