"""GPU (-m gpu): the non-quantile heads (gaussian, residual_magnitude(_l1), quantiles_l1, inn, softmax) through the C
ABI against fixtures generated from the unmodified reference (tests/golden/make_golden_heads.py) and the oracle.

Bar: bit-exact miss counts / loss tables / lambda-hat / interval endpoints for every head.  The softmax head's
lambda-independent half (softmax -> quantile planes) matches except at cumulative-probability threshold ties, which
torch's own CPU and CUDA kernels do not agree on either; everything downstream of those planes is bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from im2im_uq_b200 import _lib, rcps
    from im2im_uq_b200.calibration import calibrate_model as cm
    from im2im_uq_b200.models import heads
    from im2im_uq_b200.models.add_uncertainty import ModelWithUncertainty, add_uncertainty, _OTHER_HEADS
    from im2im_uq_b200.models.unet import UNet
    from oracle import rcps_oracle as orc
    DEV = torch.device("cuda:0")
    KIND = {"quantiles_l1": _lib.IM2IM_HEAD_QUANTILES, "inn": _lib.IM2IM_HEAD_QUANTILES,
            "residual_magnitude": _lib.IM2IM_HEAD_RESIDUAL, "residual_magnitude_l1": _lib.IM2IM_HEAD_RESIDUAL,
            "gaussian": _lib.IM2IM_HEAD_GAUSSIAN, "softmax_sets": _lib.IM2IM_HEAD_SOFTMAX_SETS}


class _Identity(torch.nn.Module):
    def forward(self, x):
        return x


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _model(g):
    _, loss_fn, set_fn = _OTHER_HEADS[g["head"]]
    return ModelWithUncertainty(_Identity(), _Identity(), loss_fn, set_fn, dict(g["config"], device="cuda:0")).to(DEV)


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("grid", ["lam_prime", "lambdas"])
def test_head_miss_counts_bit_exact_vs_reference(head_golden, grid, generic):
    g = head_golden
    key = "counts_prime" if grid == "lam_prime" else "counts_grid"
    counts, totals = rcps.miss_counts(_dev(g["scores"]), _dev(g["labels"]), _dev(g[grid]), force_generic=generic,
                                      head=KIND[g["score_head"]])
    assert np.array_equal(counts.cpu().numpy(), g[key])
    assert np.array_equal(totals.cpu().numpy(), g[key].sum(0, dtype=np.int64))


def test_head_nested_sets_vs_reference(head_golden):
    g = head_golden
    model = _model(g)
    mid = g["lam_prime"][len(g["lam_prime"]) // 2]
    for lam, lo_ref, up_ref in ((g["lhat"], g["lower_at_lhat"], g["upper_at_lhat"]), (mid, g["lower_mid"], g["upper_mid"])):
        # through the raw kernel on the score planes ...
        lo, pred, up = rcps.head_nested_sets(_dev(g["scores"]), float(lam), KIND[g["score_head"]])
        assert np.array_equal(lo.cpu().numpy(), lo_ref, equal_nan=True)
        assert np.array_equal(up.cpu().numpy(), up_ref, equal_nan=True)
        assert np.array_equal(pred.cpu().numpy(), g["pred_at_lhat"], equal_nan=True)
        fin = np.isfinite(lo_ref) & np.isfinite(up_ref)
        np.testing.assert_allclose(lo.cpu().numpy()[fin], lo_ref[fin], rtol=1e-5, atol=0)   # north_star tolerance
        # ... and through the reference-facing method (softmax: from the logits, via im2im_softmax_sets)
        lo2, pred2, up2 = model.nested_sets_from_output(_dev(g["outputs"]), torch.tensor(lam))
        if g["head"] != "softmax":
            assert np.array_equal(lo2.cpu().numpy(), lo_ref, equal_nan=True)
            assert np.array_equal(up2.cpu().numpy(), up_ref, equal_nan=True)
        else:
            same = (lo2.cpu().numpy() == lo_ref) & (up2.cpu().numpy() == up_ref)
            assert (g["threshold_margin"][~same] < 2e-6).all() and same.mean() > 0.995
    with pytest.raises(Exception, match="You have to specify lambda"):
        model.nested_sets_from_output(_dev(g["outputs"]))


def test_softmax_sets_kernel_vs_reference(head_golden):
    g = head_golden
    if g["head"] != "softmax":
        pytest.skip("softmax only")
    sets = rcps.softmax_sets(_dev(g["outputs"])).cpu().numpy()
    # bit for bit with the oracle: same operation order (division per class, double cumulative sum) and the same portable exp
    assert np.array_equal(sets, orc.softmax_sets(g["outputs"]))
    # against the reference's own planes (torch CPU: its exp and summation order round differently in the last place): only
    # pixels whose cumulative probability lies within rounding distance of 0.05 / 0.95 may differ
    same = (sets == g["softmax_sets"]).all(axis=1)
    assert (g["threshold_margin"][~same] < 2e-6).all()
    assert same.mean() > 0.995
    # values are multiples of 1/K in [0, 1]
    k = g["outputs"].shape[1]
    assert np.allclose(sets * k, np.round(sets * k), atol=1e-4) and sets.min() >= 0 and sets.max() <= 1


@pytest.mark.parametrize("K,shape", [(50, (3, 40, 36)), (64, (2, 17, 9)), (7, (5, 8, 8)), (1, (2, 4, 4)),
                                     (50, (16, 128, 128))])   # 13 M divisions through the one-reciprocal (Markstein) path
def test_softmax_sets_kernel_equals_oracle_on_random_logits(K, shape):
    """Integer result, no tolerance: planes identical to the oracle's for wide-range logits, NaN / inf rows included."""
    n, h, w = shape
    g = torch.Generator().manual_seed(K)
    logits = torch.randn(n, K, h, w, generator=g) * 6
    flat = logits.view(-1)
    idx = torch.randperm(flat.numel(), generator=g)[:24]
    for j, v in enumerate((float("nan"), float("inf"), -float("inf"), 80.0, -120.0, 0.0)):
        flat[idx[4 * j:4 * j + 4]] = v
    want = orc.softmax_sets(logits.numpy())
    got = rcps.softmax_sets(logits.to(DEV)).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("resident", [True, False])
def test_head_calibration_matches_reference(head_golden, resident):
    g = head_golden
    model = _model(g)
    out = torch.from_numpy(g["outputs"].copy())
    lab = torch.from_numpy(g["labels"].copy())
    if resident:
        out, lab = out.to(DEV), lab.to(DEV)
    cfg = dict(g["config"], device="cuda:0")
    model, table = cm.calibrate_from_outputs(model, out, lab, cfg)
    if g["head"] == "softmax":
        sets = rcps.softmax_sets(_dev(g["outputs"])).cpu().numpy()
        if not np.array_equal(sets, g["softmax_sets"]):
            ties = int((sets != g["softmax_sets"]).any(axis=1).sum())
            px = g["labels"][0].size
            assert np.abs(table.numpy() - g["calib_loss_table"]).max() <= ties / px + 1e-7
            return
    assert float(model.lhat) == float(g["lhat"])
    assert np.array_equal(table.numpy(), g["calib_loss_table"])
    assert table.device.type == "cpu" and table.dtype == torch.float32


def test_head_losses_per_image_and_miss_map(head_golden):
    g = head_golden
    model = _model(g)
    j = len(g["lam_prime"]) // 2
    lam = torch.tensor(g["lam_prime"][j])
    px = g["labels"][0].size
    ds = torch.utils.data.TensorDataset(torch.from_numpy(g["outputs"].copy()), torch.from_numpy(g["labels"].copy()))
    losses = cm.get_rcps_losses_from_outputs(model, ds, cm.fraction_missed_loss, lam, "cuda:0")
    if g["head"] != "softmax":
        assert np.array_equal(losses.numpy(), g["counts_prime"][:, j].astype(np.float32) / np.float32(px))
    mm = rcps.miss_map(_dev(g["scores"]), _dev(g["labels"]), float(lam), head=KIND[g["score_head"]])
    assert np.array_equal(mm.cpu().numpy(), orc.head_miss_map(g["scores"], g["labels"], float(lam), g["score_head"]))


@pytest.mark.parametrize("head", ["residual_magnitude", "gaussian", "softmax_sets"])
@pytest.mark.parametrize("shape", [(9, 1, 100, 100), (5, 2, 37, 41)])
def test_head_random_nasty_vs_oracle(head, shape):
    """Bigger than the fixtures, both kernels, widths of every sign / nan / inf, a grid that straddles zero."""
    n, c, h, w = shape
    g = torch.Generator().manual_seed(n * 7 + h)
    pred = torch.rand(shape, generator=g)
    sig = 0.02 + 0.1 * torch.rand(shape, generator=g)
    label = pred + sig * torch.randn(shape, generator=g)
    width = sig * (0.5 + torch.rand(shape, generator=g))
    nasty = torch.tensor([float("nan"), float("inf"), -float("inf"), 0.0, -0.0, -0.03, 1e-30, 1e30, -1e-3])
    sel = torch.randperm(pred.numel(), generator=g)[:400]
    if head == "softmax_sets":
        k = 50
        lq = torch.floor((pred - width).clamp(0, 1) * k) / k
        uq = torch.floor((pred + width).clamp(0, 1) * k) / k
        pr = torch.floor(pred * k) / k
        out = torch.stack([lq, pr, uq], dim=1)
        out[:, 0].reshape(-1)[sel[:50]] = 1.0      # lower quantile above the prediction -> relu clips the width to 0
    else:
        if head == "gaussian":
            width = width ** 2
        width.view(-1)[sel] = nasty[torch.randint(0, nasty.numel(), (400,), generator=g)]
        out = torch.stack([pred, width], dim=1)
    label.view(-1)[sel[100:140]] = float("nan")
    lams = torch.linspace(-1.5, 5.0, 333)
    want = orc.head_miss_table(out.numpy(), label.numpy(), lams.numpy(), head)
    for generic in (False, True):
        counts, totals = rcps.miss_counts(out.to(DEV), label.to(DEV), lams.to(DEV), force_generic=generic,
                                          head=KIND[head])
        assert np.array_equal(counts.cpu().numpy(), want), (head, generic)
        assert np.array_equal(totals.cpu().numpy(), want.sum(0, dtype=np.int64))


def test_head_loss_kernels_match_reference_kats():
    kats = np.load(os.path.join(GOLDEN, "head_loss_kats.npz"))
    fns = dict(gaussian=heads.gaussian_regression_loss_fn, residual_magnitude=heads.residual_magnitude_loss_fn,
               residual_magnitude_l1=heads.residual_magnitude_l1_loss_fn, quantiles_l1=heads.quantile_regression_l1_loss_fn,
               inn=heads.inn_loss_fn, softmax=heads.softmax_loss_fn)
    before = _lib.launch_count()
    for head, fn in fns.items():
        for tag in ("a", "b"):
            key = f"{head}_{tag}"
            params = dict(json.loads(str(kats[key + "_params"])), device="cuda:0")
            pred = _dev(kats[key + "_pred"]).requires_grad_(True)
            loss = fn(pred, _dev(kats[key + "_target"]), params)
            loss.backward()
            np.testing.assert_allclose(loss.item(), float(kats[key + "_loss"]), rtol=2e-6)
            np.testing.assert_allclose(pred.grad.cpu().numpy(), kats[key + "_grad"], rtol=2e-5, atol=1e-9)
    assert _lib.launch_count() - before >= 10   # the five affine heads ran on im2im_head_loss_f32


@pytest.mark.parametrize("head", ["gaussian", "residual_magnitude", "quantiles_l1", "inn", "softmax"])
def test_head_forward_on_native_engine(head):
    """eval forward: native trunk (+ native stacked head conv) vs the same modules through torch, bf16 tolerance."""
    params = dict(uncertainty_type=head, q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0, beta=0.1,
                  num_softmax=50)
    torch.manual_seed(3)
    model = add_uncertainty(UNet(1, 1), params).to(DEV)
    model.train()
    with torch.no_grad():
        model.use_native_training = False
        for _ in range(2):
            model(torch.randn(4, 1, 32, 32, device=DEV))     # move the BatchNorm statistics off their initial values
    model.eval()
    x = torch.randn(2, 1, 48, 32, device=DEV)
    with torch.no_grad():
        before = _lib.launch_count()
        y = model(x)
        assert _lib.launch_count() - before >= 20               # the tcgen05 trunk ran
        model.use_native_inference = False
        ref = model(x)
    assert y.shape == ref.shape
    scale = ref.abs().max().item()
    assert (y - ref).abs().max().item() <= 4e-2 * scale
    assert ((y - ref).norm() / ref.norm()).item() <= 2e-2
    if head == "gaussian":
        assert (y[:, 1] >= 0).all()
    if head == "residual_magnitude":
        assert (y[:, 1] >= 0).all()


@pytest.mark.parametrize("head", ["gaussian", "residual_magnitude_l1", "inn"])
def test_head_training_step_matches_autograd(head):
    """train forward/backward on the native engine vs torch fp32 autograd of the same modules.  Yardstick as in
    test_unet_train_gpu.py: the native bf16 engine must be as close to fp32 as torch's own bf16 autocast is."""
    params = dict(uncertainty_type=head, q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0, beta=0.1)

    def build():
        torch.manual_seed(5)
        m = add_uncertainty(UNet(1, 1), params).to(DEV).train()
        if head == "gaussian":
            m.last_layer.variance.bias.data.fill_(1.0)   # keep the variance off the 1e-6 clamp (1/var gradients explode)
        return m

    def run(native, autocast=False):
        model = build()
        model.use_native_training = native
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = model(x)
            out = out.float()
        else:
            out = model(x)
        loss = model.loss_fn(out, y)
        loss.backward()
        if native:
            assert "_native_train_engine" in model.__dict__
        return out.detach(), loss.item(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-30)).item()

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator(device=DEV).manual_seed(1)
        x = torch.randn(8, 1, 64, 64, device=DEV, generator=g)
        y = x + 0.3 * torch.randn(8, 1, 64, 64, device=DEV, generator=g)
        p_ref, l_ref, g_ref = run(False)
        p_ac, l_ac, g_ac = run(False, autocast=True)
        p_nat, l_nat, g_nat = run(True)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert p_nat.shape == p_ref.shape
    assert abs(l_nat - l_ref) <= max(5e-3 * abs(l_ref), 2 * abs(l_ac - l_ref))
    assert rel(p_nat, p_ref) <= max(6e-2, 1.2 * rel(p_ac, p_ref))
    names = [n for n in g_ref if g_ref[n].norm() > 1e-6 and not (n.endswith("double_conv.0.bias") or n.endswith("double_conv.3.bias"))]
    err_nat = np.array([rel(g_nat[n], g_ref[n]) for n in names])
    err_ac = np.array([rel(g_ac[n], g_ref[n]) for n in names])
    assert np.median(err_nat) <= max(0.25, 1.2 * np.median(err_ac)), (np.median(err_nat), np.median(err_ac))
    assert err_nat.max() <= max(0.8, 1.5 * err_ac.max()), (names[int(err_nat.argmax())], err_nat.max(), err_ac.max())


# ------------------------------------------------------------------------------------------- eval.py mirror
def test_get_loss_table_matches_reference_inner_loop(head_golden):
    """core/scripts/eval.py:84-126 through the drop-in: dense table at lambdas[j] == the reference's per-(batch, lambda)
    loop (fixture key counts_grid, produced with the reference's nested_sets_from_output + fraction_missed_loss)."""
    from core.scripts.eval import get_loss_table
    g = head_golden
    if g["head"] == "softmax" and not np.array_equal(rcps.softmax_sets(_dev(g["outputs"])).cpu().numpy(), g["softmax_sets"]):
        pytest.skip("threshold ties differ (covered by test_head_calibration_matches_reference)")
    model = _model(g)
    ds = torch.utils.data.TensorDataset(torch.from_numpy(g["outputs"].copy()), torch.from_numpy(g["labels"].copy()))
    table = get_loss_table(model, ds, dict(g["config"], device="cuda:0"))
    px = g["labels"][0].size
    assert table.device.type == "cpu" and table.dtype == torch.float32
    assert np.array_equal(table.numpy(), g["counts_grid"].astype(np.float32) / np.float32(px))


def test_eval_set_metrics_consistent_with_calibration(head_golden):
    from core.scripts.eval import eval_set_metrics
    g = head_golden
    model = _model(g)
    model.set_lhat(torch.tensor(g["lhat"]))
    ds = torch.utils.data.TensorDataset(torch.from_numpy(g["outputs"].copy()), torch.from_numpy(g["labels"].copy()))
    np.random.seed(0); torch.manual_seed(0)
    risk, sizes, spearman, strat, mse, spatial = eval_set_metrics(model, ds, dict(g["config"], device="cuda:0"))
    n, px = g["labels"].shape[0], g["labels"][0].size
    want = orc.head_miss_table(g["scores"], g["labels"], [float(g["lhat"])], g["score_head"])[:, 0]
    if g["head"] != "softmax":
        assert abs(float(risk) - float((want.astype(np.float32) / np.float32(px)).mean())) < 1e-6
        mm = orc.head_miss_map(g["scores"], g["labels"], float(g["lhat"]), g["score_head"])
        assert np.array_equal(spatial, (mm.astype(np.float32) / np.float32(n)).mean(axis=0))
    assert sizes.shape == (n,) and strat.shape == (4,) and spatial.shape == g["labels"].shape[2:]


# ------------------------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("head", ["residual_magnitude", "gaussian", "softmax_sets"])
def test_head_full_size_properties(head):
    """BASELINE config C2's size (1k x 320^2, L = 1000) for the other head kinds: size-independent invariants (nested
    sets are monotone in lambda, checksum of checksums, both kernels agree, shard additivity, chunked accumulation,
    idempotence) + random rows against the oracle at full L."""
    n, h, w, L = 1000, 320, 320, 1000
    g = torch.Generator(device=DEV).manual_seed(3)
    shape = (n, 1, h, w)
    pred = torch.rand(shape, generator=g, device=DEV)
    sig = 0.02 + 0.1 * torch.rand(shape, generator=g, device=DEV)
    lab = pred + sig * torch.randn(shape, generator=g, device=DEV)
    width = sig * (0.5 + torch.rand(shape, generator=g, device=DEV))
    if head == "softmax_sets":
        k = 50
        out = torch.stack([torch.floor((pred - width).clamp(0, 1) * k) / k, torch.floor(pred * k) / k,
                           torch.floor((pred + width).clamp(0, 1) * k) / k], dim=1).contiguous()
    else:
        out = torch.stack([pred, width ** 2 if head == "gaussian" else width], dim=1).contiguous()
    kind = KIND[head]
    lam_cpu = torch.linspace(0.0, 6.0, L) - (torch.linspace(0.0, 6.0, L)[1] - torch.linspace(0.0, 6.0, L)[0])
    lam = lam_cpu.to(DEV)
    px = h * w
    c, t = rcps.miss_counts(out, lab, lam, head=kind)
    assert int(c.min()) >= 0 and int(c.max()) <= px
    assert bool((c[:, 1:] <= c[:, :-1]).all())
    assert torch.equal(t, c.sum(0, dtype=torch.int64))
    cg, tg = rcps.miss_counts(out, lab, lam, force_generic=True, head=kind)
    assert torch.equal(c, cg) and torch.equal(t, tg)
    h1 = n // 3
    ca, ta = rcps.miss_counts(out[:h1], lab[:h1], lam, head=kind)
    cb, tb = rcps.miss_counts(out[h1:], lab[h1:], lam, head=kind)
    assert torch.equal(torch.cat([ca, cb]), c) and torch.equal(ta + tb, t)
    c2 = torch.zeros_like(c); t2 = torch.zeros_like(t)
    rcps.miss_counts(out[:h1], lab[:h1], lam, counts=c2[:h1], totals=t2, zero=False, head=kind)
    rcps.miss_counts(out[h1:], lab[h1:], lam, counts=c2[h1:], totals=t2, zero=False, head=kind)
    assert torch.equal(c2, c) and torch.equal(t2, t)
    c3, t3 = rcps.miss_counts(out, lab, lam, head=kind)
    assert torch.equal(c3, c) and torch.equal(t3, t)
    rows = torch.randperm(n)[:4].to(DEV)
    want = orc.head_miss_table(out[rows].cpu().numpy(), lab[rows].cpu().numpy(), lam_cpu.numpy(), head)
    assert np.array_equal(c[rows].cpu().numpy(), want)
    # interval endpoints at one lambda, whole set, against the oracle on the same rows (bit-exact)
    lo, pr, up = rcps.head_nested_sets(out[rows].contiguous(), float(lam_cpu[L // 3]), kind)
    lo_w, _, up_w = orc.head_nested_sets(out[rows].cpu().numpy(), float(lam_cpu[L // 3]), head)
    assert np.array_equal(lo.cpu().numpy(), lo_w) and np.array_equal(up.cpu().numpy(), up_w)
