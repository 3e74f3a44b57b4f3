"""CPU: the oracle (C and numpy restatements) against the fixtures generated from the unmodified reference."""
import json
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import rcps_oracle as orc


def test_c_oracle_tables_match_reference(golden):
    out, lab = golden["outputs"], golden["labels"]
    assert np.array_equal(orc.c_miss_table(out, lab, golden["lam_prime"]), golden["counts_prime"])
    assert np.array_equal(orc.c_miss_table(out, lab, golden["lambdas"]), golden["counts_grid"])


def test_numpy_oracle_matches_reference(golden):
    out, lab = golden["outputs"], golden["labels"]
    L = len(golden["lambdas"])
    for j in sorted({0, 1, L // 3, L // 2, L - 2, L - 1}):
        assert np.array_equal(orc.np_miss_counts(out, lab, golden["lam_prime"][j]), golden["counts_prime"][:, j])
        assert np.array_equal(orc.c_miss_counts(out, lab, golden["lam_prime"][j]), golden["counts_prime"][:, j])


def test_loss_is_count_over_pixels(golden):
    px = np.float32(np.prod(golden["outputs"].shape[2:]))
    assert np.array_equal(golden["counts_prime"].astype(np.float32) / px, golden["dense_prime"])
    assert np.array_equal(golden["counts_grid"].astype(np.float32) / px, golden["dense_grid"])


def test_nested_sets_match_reference(golden):
    lo, p, up = orc.c_nested_sets(golden["outputs"], float(golden["lhat"]))
    assert np.array_equal(lo, golden["lower_at_lhat"], equal_nan=True)
    assert np.array_equal(up, golden["upper_at_lhat"], equal_nan=True)
    lo2, p2, up2 = orc.np_nested_sets(golden["outputs"], golden["lhat"])
    assert np.array_equal(lo2, golden["lower_at_lhat"], equal_nan=True)
    assert np.array_equal(up2, golden["upper_at_lhat"], equal_nan=True)


def test_sweep_restatement_matches_reference(golden):
    cfg = golden["config"]
    lhat, stop, table = orc.calibrate_sweep(golden["outputs"], golden["labels"], cfg["minimum_lambda"],
                                            cfg["maximum_lambda"], cfg["num_lambdas"], cfg["alpha"], cfg["delta"])
    assert stop == int(golden["stop_idx"])
    assert np.float32(lhat.numpy()) == golden["lhat"]
    assert np.array_equal(table.numpy(), golden["calib_loss_table"])


def test_miss_map_consistent_with_counts(golden):
    lam = float(golden["lam_prime"][len(golden["lam_prime"]) // 2])
    m = orc.c_miss_map(golden["outputs"], golden["labels"], lam)
    assert m.sum() == golden["counts_prime"][:, len(golden["lam_prime"]) // 2].sum()


def test_hb_known_answers():
    kats = json.load(open(os.path.join(GOLDEN, "hb_mu_plus_kats.json")))
    assert len(kats) > 300
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in kats:
            assert orc.hb_mu_plus(k["muhat"], k["n"], k["delta"]) == k["value"], k


def test_survey_probe_kats():
    # values recorded from the reference during the survey
    assert orc.hb_mu_plus(0.1, 10000, 0.1, 1000) == 0.10551758004098837
    assert orc.hb_mu_plus(0.09, 1000, 0.1) == 0.10774366846869704
    assert orc.hb_mu_plus(0.05, 1000, 0.1) == 0.0640138403022351
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert orc.hb_mu_plus(0.0, 1000, 0.1) == 1.0
    assert orc.hb_mu_plus(1.0, 100, 0.1) == 1
