"""GPU (-m gpu): parity of the sm_100a RCPS path, called through the C ABI, against the oracle / reference fixtures.

Bar: bit-exact for miss counts, totals, lambda-hat index and the fp32 loss table; interval endpoints are also produced
bit-exactly, and additionally asserted within the north_star's 1e-5 relative tolerance.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, synth_scores

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from im2im_uq_b200 import _lib, rcps
    from im2im_uq_b200.calibration import calibrate_model as cm
    from im2im_uq_b200.calibration import sweep
    from im2im_uq_b200.models.add_uncertainty import ModelWithUncertainty
    from im2im_uq_b200.models.quantile_layer import (quantile_regression_loss_fn,
                                                     quantile_regression_nested_sets_from_output)
    from oracle import rcps_oracle as orc
    DEV = torch.device("cuda:0")

ENDPOINT_RTOL = 1e-5  # BASELINE.json north_star tolerance for float interval endpoints


class _Identity(torch.nn.Module):
    def forward(self, x):
        return x


def _identity_model(params):
    return ModelWithUncertainty(_Identity(), _Identity(), quantile_regression_loss_fn,
                                quantile_regression_nested_sets_from_output, params)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ------------------------------------------------------------------------------------------- golden parity
@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("grid", ["lam_prime", "lambdas"])
def test_miss_counts_bit_exact_vs_reference(golden, grid, generic):
    key = "counts_prime" if grid == "lam_prime" else "counts_grid"
    counts, totals = rcps.miss_counts(_dev(golden["outputs"]), _dev(golden["labels"]), _dev(golden[grid]),
                                      force_generic=generic)
    assert np.array_equal(counts.cpu().numpy(), golden[key])
    assert np.array_equal(totals.cpu().numpy(), golden[key].sum(0, dtype=np.int64))
    px = int(np.prod(golden["outputs"].shape[2:]))
    dense = "dense_prime" if grid == "lam_prime" else "dense_grid"
    assert np.array_equal(rcps.loss_table(counts, px).cpu().numpy(), golden[dense])  # get_loss_table format


def test_nested_sets_vs_reference(golden):
    out = _dev(golden["outputs"])
    lo, pred, up = rcps.quantile_nested_sets(out, float(golden["lhat"]))
    for got, want in ((lo, golden["lower_at_lhat"]), (up, golden["upper_at_lhat"]), (pred, golden["pred_at_lhat"])):
        got = got.cpu().numpy()
        assert np.array_equal(got, want, equal_nan=True)
        fin = np.isfinite(want)
        np.testing.assert_allclose(got[fin], want[fin], rtol=ENDPOINT_RTOL, atol=0)
    # reference side effect: `output` planes 0/2 clamped in place (quantile_layer.py:39-40)
    o = golden["outputs"]
    with np.errstate(all="ignore"):
        assert np.array_equal(out[:, 0].cpu().numpy(), np.minimum(o[:, 0], o[:, 1] - np.float32(1e-6)), equal_nan=True)
        assert np.array_equal(out[:, 2].cpu().numpy(), np.maximum(o[:, 2], o[:, 1] + np.float32(1e-6)), equal_nan=True)
    assert pred.data_ptr() == out[:, 1].data_ptr()  # prediction is a view, like the reference


def test_model_nested_sets_from_output(golden):
    model = _identity_model(golden["config"]).to(DEV)
    with pytest.raises(Exception, match="You have to specify lambda"):
        model.nested_sets_from_output(_dev(golden["outputs"]))
    model.set_lhat(torch.tensor(golden["lhat"]))
    lo, pred, up = model.nested_sets_from_output(_dev(golden["outputs"]))
    assert np.array_equal(lo.cpu().numpy(), golden["lower_at_lhat"], equal_nan=True)
    assert np.array_equal(up.cpu().numpy(), golden["upper_at_lhat"], equal_nan=True)
    lo2, _, up2 = model.nested_sets((_dev(golden["outputs"]),))  # through forward (identity trunk/head)
    assert torch.equal(torch.nan_to_num(lo2), torch.nan_to_num(lo)) and torch.equal(torch.nan_to_num(up2), torch.nan_to_num(up))


def test_fraction_missed_loss_and_miss_map(golden):
    j = len(golden["lam_prime"]) // 2
    lam = float(golden["lam_prime"][j])
    out, lab = _dev(golden["outputs"]), _dev(golden["labels"])
    sets = rcps.quantile_nested_sets(out.clone(), lam, write_back_clamp=False)
    loss = cm.fraction_missed_loss(sets, lab)
    assert np.array_equal(loss.cpu().numpy(), golden["dense_prime"][:, j])
    mm = rcps.miss_map(out, lab, lam)
    assert np.array_equal(mm.cpu().numpy(), orc.c_miss_map(golden["outputs"], golden["labels"], lam))


@pytest.mark.parametrize("resident", ["cuda", "cpu"])
def test_calibrate_from_outputs_matches_reference(golden, resident):
    cfg = dict(golden["config"], device="cuda:0")
    model = _identity_model(cfg)
    out = torch.from_numpy(golden["outputs"]); lab = torch.from_numpy(golden["labels"])
    if resident == "cuda":
        out, lab = out.to(DEV), lab.to(DEV)
    stats = {}
    model, table = cm.calibrate_from_outputs(model, out, lab, cfg, stats=stats)
    assert table.device.type == "cpu" and table.dtype == torch.float32
    assert np.array_equal(table.numpy(), golden["calib_loss_table"])
    assert np.float32(model.lhat.numpy()) == golden["lhat"] and model.lhat.dim() == 0
    assert stats["replayed_columns"] <= 3


def test_calibrate_model_drop_in(golden):
    """The reference's own entry point: calibrate_model(model, dataset, config) -> (model, table)."""
    cfg = dict(golden["config"], device="cuda")
    model = _identity_model(cfg)
    ds = torch.utils.data.TensorDataset(torch.from_numpy(golden["outputs"]), torch.from_numpy(golden["labels"]))
    model2, table = cm.calibrate_model(model, ds, cfg)
    assert model2 is model
    assert np.array_equal(table.numpy(), golden["calib_loss_table"])
    assert np.float32(model.lhat.numpy()) == golden["lhat"]


def test_get_rcps_losses_from_outputs(golden):
    cfg = dict(golden["config"], device="cuda:0")
    model = _identity_model(cfg)
    ds = torch.utils.data.TensorDataset(torch.from_numpy(golden["outputs"]), torch.from_numpy(golden["labels"]))
    for j in (0, len(golden["lambdas"]) // 2, len(golden["lambdas"]) - 1):
        lam = torch.tensor(golden["lam_prime"][j])
        losses = cm.get_rcps_losses_from_outputs(model, ds, cm.fraction_missed_loss, lam, "cuda:0")
        assert losses.device.type == "cpu" and np.array_equal(losses.numpy(), golden["dense_prime"][:, j])
        # generic (non-fused) route: any user loss on nested sets, still on the GPU
        generic = cm.get_rcps_losses_from_outputs(model, ds, lambda s, y: cm.fraction_missed_loss(s, y), lam, "cuda:0")
        assert np.array_equal(generic.numpy(), golden["dense_prime"][:, j])
    with pytest.raises(Exception, match="You have to specify lambda"):
        cm.get_rcps_losses_from_outputs(model, ds, cm.fraction_missed_loss, None, "cuda:0")


# get_rcps_metrics_from_outputs: see tests/test_metrics_golden.py (fixtures produced by the reference's own function)


def test_empty_and_degenerate_sizes():
    lam = torch.linspace(0, 2, 5, device=DEV)
    c, t = rcps.miss_counts(torch.zeros(0, 3, 1, 4, 4, device=DEV), torch.zeros(0, 1, 4, 4, device=DEV), lam)
    assert tuple(c.shape) == (0, 5) and t.tolist() == [0] * 5
    out, lab = synth_scores(1, 3, 1, 1, 1, device=DEV)  # one pixel per image
    c, t = rcps.miss_counts(out, lab, lam)
    assert np.array_equal(c.cpu().numpy(), orc.c_miss_table(out.cpu().numpy(), lab.cpu().numpy(), lam.cpu().numpy()))


@pytest.mark.parametrize("shape", [(5, 1, 63, 65), (3, 2, 17, 3), (2, 1, 1, 2049), (9, 1, 64, 64), (4, 3, 32, 20)])
def test_ragged_shapes_and_partial_tiles(shape):
    out, lab = synth_scores(7, *shape, device=DEV)
    lam = (torch.linspace(0, 6, 97) - 0.0625).to(DEV)
    want = orc.c_miss_table(out.cpu().numpy(), lab.cpu().numpy(), lam.cpu().numpy())
    for generic in (False, True):
        c, t = rcps.miss_counts(out, lab, lam, force_generic=generic)
        assert np.array_equal(c.cpu().numpy(), want)
        assert np.array_equal(t.cpu().numpy(), want.sum(0, dtype=np.int64))


def test_unaligned_and_strided_inputs():
    big_o, big_l = synth_scores(3, 12, 1, 16, 16, device=DEV)
    lam = torch.linspace(0, 4, 50, device=DEV)
    want = orc.c_miss_table(big_o.cpu().numpy(), big_l.cpu().numpy(), lam.cpu().numpy())
    # every second image: stride 2*3*px, still contiguous inside an image
    c, _ = rcps.miss_counts(big_o[::2], big_l[::2], lam)
    assert np.array_equal(c.cpu().numpy(), want[::2])
    # misaligned base pointers (offset by one float) -> the scalar-load kernel is selected automatically
    flat_o = torch.empty(big_o.numel() + 1, device=DEV); flat_l = torch.empty(big_l.numel() + 1, device=DEV)
    o2 = flat_o[1:].view_as(big_o); o2.copy_(big_o)
    l2 = flat_l[1:].view_as(big_l); l2.copy_(big_l)
    assert o2.data_ptr() % 16 != 0
    c, _ = rcps.miss_counts(o2, l2, lam)
    assert np.array_equal(c.cpu().numpy(), want)
    # separate (non-packed) channel-last style input is made contiguous by the wrapper
    perm = big_o.permute(0, 1, 2, 4, 3).contiguous().permute(0, 1, 2, 4, 3)
    c, _ = rcps.miss_counts(perm, big_l, lam)
    assert np.array_equal(c.cpu().numpy(), want)


def test_lambda_count_limits_and_irregular_grids():
    out, lab = synth_scores(11, 6, 1, 32, 32, device=DEV)
    o_np, l_np = out.cpu().numpy(), lab.cpu().numpy()
    for lam in (torch.tensor([1.25]), torch.tensor([-3.0, -1.0, -0.5]), torch.sort(torch.rand(300) * 5)[0],
                torch.cat([torch.zeros(5), torch.ones(5), torch.full((5,), 2.0)]), torch.linspace(0, 6, 8192),
                torch.linspace(0, 6, 6500)):
        c, t = rcps.miss_counts(out, lab, lam.to(DEV))
        assert np.array_equal(c.cpu().numpy(), orc.c_miss_table(o_np, l_np, lam.numpy())), lam.shape
    with pytest.raises(_lib.Im2ImError, match="n_lambdas"):
        rcps.miss_counts(out, lab, torch.linspace(0, 6, 8193, device=DEV))


def test_device_decide_agrees_with_host_screen(golden):
    """im2im_rcps_decide == sweep._classify + scan on the same totals (decided stop or the same first unsure column)."""
    cfg = golden["config"]
    px = int(np.prod(golden["outputs"].shape[2:]))
    n = golden["outputs"].shape[0]
    totals = torch.from_numpy(golden["counts_prime"].sum(0, dtype=np.int64)).to(DEV)
    stop, decided = cm.device_decide_fn(px, cfg)(totals, n)
    verdict = sweep._classify(totals.cpu().numpy(), n, px, cfg["alpha"], cfg["delta"])
    nonfalse = np.nonzero(verdict >= 0)[0]
    if nonfalse.size == 0:
        assert decided and stop == -1
    elif verdict[nonfalse[-1]] > 0:
        assert decided and stop == int(nonfalse[-1]) == int(golden["stop_idx"])
    else:
        assert not decided


def test_descending_grid_and_batch_of_65():
    g = load_golden("batch65")  # N % 64 == 2; also run N % 64 == 1 (the reference itself raises there)
    for n in (66, 65):
        out, lab = g["outputs"][:n], g["labels"][:n]
        cfg = dict(g["config"], device="cuda:0")
        model, table = cm.calibrate_from_outputs(_identity_model(cfg), _dev(out), _dev(lab), cfg)
        lhat, stop, ref_table = orc.calibrate_sweep(out, lab, cfg["minimum_lambda"], cfg["maximum_lambda"],
                                                    cfg["num_lambdas"], cfg["alpha"], cfg["delta"])
        assert torch.equal(model.lhat, lhat) and torch.equal(table, ref_table)
    cfg = dict(g["config"], device="cuda:0", minimum_lambda=4.0, maximum_lambda=0.0)  # descending grid
    model, table = cm.calibrate_from_outputs(_identity_model(cfg), _dev(g["outputs"]), _dev(g["labels"]), cfg)
    lhat, stop, ref_table = orc.calibrate_sweep(g["outputs"], g["labels"], 4.0, 0.0, cfg["num_lambdas"], cfg["alpha"],
                                                cfg["delta"])
    assert torch.equal(model.lhat, lhat) and torch.equal(table, ref_table)


# ------------------------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("n,h,w,grid", [(1000, 320, 320, (0.0, 6.0, 1000)), (192, 512, 512, (7.0, 10.0, 100))])
def test_full_size_properties(n, h, w, grid):
    """BASELINE configs C2 (fastmri 1k x 320^2, L=1000) and C4's image size/grid: size-independent invariants +
    a random-row check against the oracle."""
    out, lab = synth_scores(0, n, 1, h, w, device=DEV, noise=1.0 if grid[0] == 0.0 else 6.0)
    cfg = dict(uncertainty_type="quantiles", minimum_lambda=grid[0], maximum_lambda=grid[1], num_lambdas=grid[2])
    lambdas, dl, lam_prime, _ = sweep.lambda_grid(cfg)
    lam = lam_prime.to(DEV)
    px = h * w
    c, t = rcps.miss_counts(out, lab, lam)
    assert int(c.min()) >= 0 and int(c.max()) <= px
    assert bool((c[:, 1:] <= c[:, :-1]).all())                       # nested sets: misses never increase with lambda
    assert torch.equal(t, c.sum(0, dtype=torch.int64))                # checksum of checksums
    cg, tg = rcps.miss_counts(out, lab, lam, force_generic=True)
    assert torch.equal(c, cg) and torch.equal(t, tg)                  # both kernels agree bit for bit
    # shard additivity (the multi-GPU decomposition): halves reproduce the rows, totals add up
    h1 = n // 3
    ca, ta = rcps.miss_counts(out[:h1], lab[:h1], lam)
    cb, tb = rcps.miss_counts(out[h1:], lab[h1:], lam)
    assert torch.equal(torch.cat([ca, cb]), c) and torch.equal(ta + tb, t)
    # chunked accumulation into preallocated outputs (how CPU-resident scores are fed)
    c2 = torch.zeros_like(c); t2 = torch.zeros_like(t)
    rcps.miss_counts(out[:h1], lab[:h1], lam, counts=c2[:h1], totals=t2, zero=False)
    rcps.miss_counts(out[h1:], lab[h1:], lam, counts=c2[h1:], totals=t2, zero=False)
    assert torch.equal(c2, c) and torch.equal(t2, t)
    # permutation of images permutes rows
    perm = torch.randperm(n, device=DEV)
    cp, tp = rcps.miss_counts(out[perm].contiguous(), lab[perm].contiguous(), lam)
    assert torch.equal(cp, c[perm]) and torch.equal(tp, t)
    # idempotence: same call, same bits
    c3, t3 = rcps.miss_counts(out, lab, lam)
    assert torch.equal(c3, c) and torch.equal(t3, t)
    # random rows against the oracle at full L
    rows = torch.randperm(n)[:6]
    want = orc.c_miss_table(out[rows.to(DEV)].cpu().numpy(), lab[rows.to(DEV)].cpu().numpy(), lam_prime.numpy())
    assert np.array_equal(c[rows.to(DEV)].cpu().numpy(), want)


def test_full_size_calibration_decision_matches_linear_scan():
    """C2-sized sweep: the screened decision equals a literal reverse linear scan over the same counts."""
    n = 1000
    out, lab = synth_scores(0, n, 1, 320, 320, device=DEV)
    cfg = dict(uncertainty_type="quantiles", minimum_lambda=0.0, maximum_lambda=6.0, num_lambdas=1000, alpha=0.1,
               delta=0.1, device="cuda:0", dataset="synthetic", rcps_loss="fraction_missed")
    stats = {}
    lhat, stop, counts, visited = cm.rcps_sweep(out, lab, cfg, stats=stats)
    px = 320 * 320
    table = counts.float().cpu() / float(px)
    from im2im_uq_b200.calibration.bounds import HB_mu_plus
    ref_stop = -1
    for j in reversed(range(1000)):
        rhat = table[:, j].mean()
        if rhat >= 0.1 or HB_mu_plus(rhat.item(), n, 0.1) > 0.1:
            ref_stop = j
            break
    assert stop == ref_stop and 300 < stop < 500 and stats["replayed_columns"] <= 3
    assert stats.get("decided_on_device") or stats["replayed_columns"] >= 1  # device screen, host replay only in the band
    lambdas = torch.linspace(0.0, 6.0, 1000)
    assert torch.equal(lhat, lambdas[stop]) and int(visited.sum()) == 1000 - stop


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_sweep():
    import subprocess, sys, os
    from conftest import ROOT
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "nccl_sweep_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "NCCL_SWEEP_OK" in res.stdout


@pytest.mark.parametrize("fused", [True, False])
def test_rcps_graph_replay_matches_reference(golden, fused):
    """The CUDA-graph plan (capture once, replay) gives the same lhat / table as the reference, replay after replay -
    as ONE fused launch per calibration (im2im_rcps_calibrate_fused) and as the multi-launch sequence."""
    cfg = dict(golden["config"], device="cuda:0")
    out, lab = _dev(golden["outputs"]), _dev(golden["labels"])
    plan = cm.RcpsGraph(out, lab, cfg, fused=fused)
    aligned = golden["labels"][0].size % 4 == 0
    assert plan.fused == (fused and aligned)
    assert plan.kernels_per_replay == (1 if plan.fused else 3)
    for _ in range(3):
        lhat, stop, decided = plan.run()
        if not decided:
            lhat, stop = plan.replay_on_host()
        assert stop == int(golden["stop_idx"]) and np.float32(lhat.numpy()) == golden["lhat"]
        assert np.array_equal(plan.table.cpu().numpy(), golden["calib_loss_table"])
        assert np.array_equal(plan.counts.cpu().numpy(), golden["counts_prime"])
    # new scores in the same buffers -> new answer from the same graph
    out.copy_(out.flip(0)); lab.copy_(lab.flip(0))
    lhat2, stop2, decided2 = plan.run()
    if not decided2:
        lhat2, stop2 = plan.replay_on_host()
    assert stop2 == int(golden["stop_idx"])          # the stopping rule is permutation invariant
    assert np.array_equal(plan.counts.cpu().numpy(), golden["counts_prime"][::-1])
    plan.close()


@pytest.mark.parametrize("head", ["quantiles", "residual_magnitude", "gaussian", "softmax_sets"])
@pytest.mark.parametrize("shape,L", [((300, 64, 64), 1000), ((20, 320, 320), 1000), ((400, 96, 96), 100),
                                     ((1000, 32, 32), 333), ((7, 640, 640), 1000), ((149, 128, 128), 64)])
def test_fused_single_launch_equals_multi_launch(shape, L, head):
    """One launch (no memset, head-partial rows between neighbouring blocks, last-block decision, in-kernel loss table,
    result through mapped pinned memory) against the separate kernels, on shapes where images straddle thread blocks
    (2 / 4.5 / 8 tiles per image), where the grid shrinks to one block per image, and where blocks own whole images only.
    Garbage in the output buffers beforehand proves that nothing relies on a memset."""
    if head != "quantiles" and shape[0] * shape[1] * shape[2] > 300 * 64 * 64 * 2:
        pytest.skip("other head kinds: small shapes only (same code path, templated)")
    n, h, w = shape
    out, lab = synth_scores(5, n, 1, h, w, device=DEV)
    kind = {"quantiles": _lib.IM2IM_HEAD_QUANTILES, "residual_magnitude": _lib.IM2IM_HEAD_RESIDUAL,
            "gaussian": _lib.IM2IM_HEAD_GAUSSIAN, "softmax_sets": _lib.IM2IM_HEAD_SOFTMAX_SETS}[head]
    if head in ("residual_magnitude", "gaussian"):
        width = (out[:, 2] - out[:, 1])
        out = torch.stack([out[:, 1], width if head == "residual_magnitude" else width * width], dim=1).contiguous()
    cfg = dict(uncertainty_type="quantiles", minimum_lambda=0.0, maximum_lambda=6.0, num_lambdas=L, alpha=0.1,
               delta=0.1, device="cuda:0", dataset="synthetic", rcps_loss="fraction_missed")
    ref = cm.RcpsGraph(out, lab, cfg, head=kind, fused=False)
    plan = cm.RcpsGraph(out, lab, cfg, head=kind, fused=True)
    assert plan.fused and plan.kernels_per_replay == 1 and not ref.fused
    want = ref.run()
    torch.cuda.synchronize()
    for rep in range(3):
        plan.counts.fill_(-12345); plan.table.fill_(float("nan")); plan.totals.fill_(-1)
        got = plan.run()
        torch.cuda.synchronize()
        assert (got[1], got[2]) == (want[1], want[2]) and torch.equal(got[0], want[0])
        assert torch.equal(plan.counts, ref.counts)
        assert torch.equal(plan.totals, ref.totals)
        assert torch.equal(plan.table, ref.table)
        assert torch.equal(plan.result, ref.result)
    if head == "quantiles" and n * h * w <= 300 * 64 * 64:
        _, _, lam_prime, _ = sweep.lambda_grid(cfg)
        assert np.array_equal(plan.counts.cpu().numpy(), orc.c_miss_table(out.cpu().numpy(), lab.cpu().numpy(), lam_prime.numpy()))
    # new contents, same buffers: the workspace cleaned itself
    lab.add_(0.01)
    want = ref.run(); got = plan.run()
    torch.cuda.synchronize()
    assert got[1] == want[1] and torch.equal(plan.counts, ref.counts) and torch.equal(plan.table, ref.table)
    plan.close(); ref.close()


@pytest.mark.parametrize("n,side,lam_min,lam_max,L,noise", [(1000, 320, 0.0, 6.0, 1000, 1.0), (250, 640, 0.0, 6.0, 1000, 1.0),
                                                            (400, 512, 7.0, 10.0, 100, 4.6)])
def test_fused_single_launch_at_baseline_image_sizes(n, side, lam_min, lam_max, L, noise):
    """BASELINE configs C2 / C5 / C4 image sizes and grids (per-GPU slices of them): the one-launch calibration against the
    separate kernels - counts, totals, table, decision - and the size-independent properties of the counts."""
    out, lab = synth_scores(11, n, 1, side, side, device=DEV, noise=noise)
    cfg = dict(uncertainty_type="quantiles", minimum_lambda=lam_min, maximum_lambda=lam_max, num_lambdas=L, alpha=0.1,
               delta=0.1, device="cuda:0", dataset="synthetic", rcps_loss="fraction_missed")
    ref = cm.RcpsGraph(out, lab, cfg, fused=False)
    plan = cm.RcpsGraph(out, lab, cfg, fused=True)
    assert plan.fused and plan.kernels_per_replay == 1
    want, got = ref.run(), plan.run()
    torch.cuda.synchronize()
    assert (got[1], got[2]) == (want[1], want[2])
    assert torch.equal(plan.counts, ref.counts) and torch.equal(plan.totals, ref.totals) and torch.equal(plan.table, ref.table)
    c = plan.counts
    assert bool((c[:, 1:] <= c[:, :-1]).all())                       # misses are non-increasing in lambda
    assert int(c.max()) <= side * side and int(c.min()) >= 0
    assert torch.equal(plan.totals, c.sum(0, dtype=torch.int64))     # checksum of checksums
    if got[2]:
        first = got[1] if got[1] >= 0 else 0
        assert float(plan.table[:, :first].abs().max()) == 0.0 if first > 0 else True
        # IEEE division on the CPU (torch's CUDA `tensor / scalar` multiplies by a rounded reciprocal: 1 ulp apart for a
        # pixel count that is not a power of two; the reference's CPU table - and ours - is count / px correctly rounded)
        assert torch.equal(plan.table[:, first:].cpu(), c[:, first:].cpu().float() / float(side * side))
    plan.close(); ref.close()


def test_fused_rejects_what_it_cannot_take():
    lib = _lib.load()
    out, lab = synth_scores(1, 4, 1, 9, 7, device=DEV)          # 63 values per image: not a multiple of 4
    cfg = dict(uncertainty_type="quantiles", minimum_lambda=0.0, maximum_lambda=6.0, num_lambdas=50, alpha=0.1,
               delta=0.1, device="cuda:0", dataset="synthetic", rcps_loss="fraction_missed")
    plan = cm.RcpsGraph(out, lab, cfg)
    assert not plan.fused and plan.kernels_per_replay == 3      # fell back to the separate kernels
    lhat, stop, decided = plan.run()
    plan.close()
    assert lib.im2im_rcps_fused_workspace_bytes(0) == 0 and lib.im2im_rcps_fused_workspace_bytes(1000) > 8000
    assert lib.im2im_host_wait_flag(None, 1, 10) == _lib_einval()


def _lib_einval():
    return -22


def test_decide_p2p_single_rank_equals_decide(golden):
    """im2im_rcps_decide_p2p with world = 1 (the mailbox and flags are this GPU's own buffers): same decision and totals
    as im2im_rcps_decide, call after call (the epoch advances on the device).  The multi-rank exchange itself is covered
    by tests/nccl_sweep_worker.py on a 2-GPU box."""
    lib = _lib.load()
    cfg = golden["config"]
    n, L = golden["outputs"].shape[0], len(golden["lam_prime"])
    px = int(np.prod(golden["labels"].shape[1:]))
    _, totals = rcps.miss_counts(_dev(golden["outputs"]), _dev(golden["labels"]), _dev(golden["lam_prime"]))
    n_px, gamma, alpha32, r_lo, r_hi, slack = sweep.screening_constants(n, px, cfg["alpha"], cfg["delta"])
    want = torch.empty(4, dtype=torch.int32, device=DEV)
    st = torch.cuda.current_stream(DEV).cuda_stream
    _lib.check(lib.im2im_rcps_decide(totals.data_ptr(), L, n_px, gamma, alpha32, r_lo, r_hi, slack, want.data_ptr(), st))
    mailbox = torch.full((2 * 1 * L,), -7, dtype=torch.int64, device=DEV)      # garbage: must be overwritten before use
    flags = torch.zeros(32, dtype=torch.int32, device=DEV)
    epoch = torch.zeros(1, dtype=torch.int32, device=DEV)
    mail_ptrs = torch.tensor([mailbox.data_ptr()], dtype=torch.int64, device=DEV)
    flag_ptrs = torch.tensor([flags.data_ptr()], dtype=torch.int64, device=DEV)
    for call in range(1, 4):
        got = torch.full((4,), -9, dtype=torch.int32, device=DEV)
        out = torch.zeros(L, dtype=torch.int64, device=DEV)
        _lib.check(lib.im2im_rcps_decide_p2p(totals.data_ptr(), mail_ptrs.data_ptr(), flag_ptrs.data_ptr(), epoch.data_ptr(),
                                             0, 1, L, n_px, gamma, alpha32, r_lo, r_hi, slack, out.data_ptr(),
                                             got.data_ptr(), st), "im2im_rcps_decide_p2p")
        assert torch.equal(got, want) and torch.equal(out, totals)
        assert int(epoch) == call and int(flags[0]) == call
