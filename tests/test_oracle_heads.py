"""CPU: the oracle's restatement of the non-quantile heads against fixtures generated from the unmodified reference
(tests/golden/make_golden_heads.py): gaussian, residual_magnitude(_l1), quantiles_l1, inn, softmax."""
import numpy as np

from oracle import rcps_oracle as orc


def test_head_tables_match_reference(head_golden):
    g = head_golden
    assert np.array_equal(orc.head_miss_table(g["scores"], g["labels"], g["lam_prime"], g["score_head"]), g["counts_prime"])
    assert np.array_equal(orc.head_miss_table(g["scores"], g["labels"], g["lambdas"], g["score_head"]), g["counts_grid"])


def test_head_nested_sets_match_reference(head_golden):
    g = head_golden
    mid = g["lam_prime"][len(g["lam_prime"]) // 2]
    for lam, lo_ref, up_ref in ((g["lhat"], g["lower_at_lhat"], g["upper_at_lhat"]), (mid, g["lower_mid"], g["upper_mid"])):
        lo, p, up = orc.head_nested_sets(g["scores"], float(lam), g["score_head"])
        assert np.array_equal(lo, lo_ref, equal_nan=True) and np.array_equal(up, up_ref, equal_nan=True)
        assert np.array_equal(p, g["pred_at_lhat"], equal_nan=True)
        lo2, _, up2 = orc.np_head_nested_sets(g["scores"], lam, g["score_head"])
        assert np.array_equal(lo2, lo_ref, equal_nan=True) and np.array_equal(up2, up_ref, equal_nan=True)


def test_head_sweep_matches_reference(head_golden):
    g, cfg = head_golden, head_golden["config"]
    lhat, stop, table = orc.calibrate_sweep(g["scores"], g["labels"], cfg["minimum_lambda"], cfg["maximum_lambda"],
                                            cfg["num_lambdas"], cfg["alpha"], cfg["delta"], head=g["score_head"])
    assert stop == int(g["stop_idx"])
    assert np.float32(lhat.numpy()) == g["lhat"]
    assert np.array_equal(table.numpy(), g["calib_loss_table"])


def test_head_miss_map_sums_to_counts(head_golden):
    g = head_golden
    j = len(g["lam_prime"]) // 2
    m = orc.head_miss_map(g["scores"], g["labels"], float(g["lam_prime"][j]), g["score_head"])
    assert m.sum() == g["counts_prime"][:, j].sum()


def test_softmax_sets_match_reference_up_to_threshold_ties(head_golden):
    g = head_golden
    if g["head"] != "softmax":
        return
    sets = orc.softmax_sets(g["outputs"])
    same = (sets == g["softmax_sets"]).all(axis=1)
    # the only admissible differences: a cumulative probability within rounding distance of 0.05 / 0.95
    assert (g["threshold_margin"][~same] < 2e-6).all()
    assert same.mean() > 0.995


def test_c_and_numpy_restatements_agree_on_random_nasty_inputs():
    """Two independent restatements (C, op-by-op; numpy array ops) of every head's chain agree bit for bit on inputs with
    NaN / inf / signed zeros / negative widths / huge and tiny magnitudes, over grids that straddle zero."""
    rng = np.random.default_rng(123)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, -0.07, 1e-30, 1e30, -1e-3, 0.5], dtype=np.float32)
    for head, planes in (("residual_magnitude", 2), ("gaussian", 2), ("softmax_sets", 3), ("quantiles_l1", 3)):
        for trial in range(6):
            shape = (5, planes, 2, 7, 9)
            out = rng.random(shape, dtype=np.float32)
            if head == "softmax_sets":
                out = np.floor(out * 50) / np.float32(50)
            lab = rng.random((5, 2, 7, 9), dtype=np.float32)
            idx = rng.integers(0, out.size, 60)
            out.reshape(-1)[idx] = special[rng.integers(0, special.size, 60)]
            idx = rng.integers(0, lab.size, 20)
            lab.reshape(-1)[idx] = special[rng.integers(0, special.size, 20)]
            lams = np.linspace(-1.0 - trial, 4.0 + trial, 17).astype(np.float32)
            table = orc.head_miss_table(out, lab, lams, head)
            for j, lam in enumerate(lams):
                counts = orc.np_fraction_missed(orc.np_head_nested_sets(out, lam, head), lab)[1]
                assert np.array_equal(counts, table[:, j]), (head, trial, j)
                lo_c, _, up_c = orc.head_nested_sets(out, float(lam), head)
                lo_n, _, up_n = orc.np_head_nested_sets(out, lam, head)
                assert np.array_equal(lo_c, lo_n, equal_nan=True) and np.array_equal(up_c, up_n, equal_nan=True)
