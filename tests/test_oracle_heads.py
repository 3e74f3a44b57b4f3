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
