"""GPU (-m gpu): streaming calibration - the quantile head's convolution books every pixel's RCPS rank in its own
epilogue (im2im_head_conv3x3_tc_hist) so that calibrate_model never writes the (N, 3, C, H, W) head tensor
(reference: calibrate_model.py:106-136 parks it on the CPU and re-reads it once per lambda).

Bar: integer results - the per-image miss counts, their totals, lhat's index and the fp32 loss table are BIT-IDENTICAL
to the two-stage path (head tensor materialised -> im2im_rcps_miss_counts), which the other GPU tests pin to the
reference's fixtures and to the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _head_inputs(B, H, W, seed, c_mid=32):
    from im2im_uq_b200.conv import pack_conv_weight, pad_head_weight
    g = torch.Generator().manual_seed(seed)
    m = torch.zeros(B, H, W, 64)
    m[..., :c_mid] = torch.randn(B, H, W, c_mid, generator=g)
    hw = torch.randn(3, c_mid, 3, 3, generator=g) * 0.08
    hb = torch.tensor([-0.4, 0.0, 0.4]) + 0.05 * torch.randn(3, generator=g)
    packed = pack_conv_weight(pad_head_weight(hw.to(DEV)))
    return m.to(DEV).to(torch.bfloat16).contiguous(), packed, hb.to(DEV), g


@pytest.mark.parametrize("B,H,W,L,lam_hi", [(3, 32, 48, 1000, 6.0), (5, 64, 64, 257, 3.0), (400, 16, 8, 100, 2.0),
                                            (2, 320, 320, 1000, 6.0), (7, 48, 40, 4000, 1.5), (1, 16, 8, 1, 1.0)])
def test_head_histogram_equals_miss_counts_of_the_materialised_head(B, H, W, L, lam_hi):
    from im2im_uq_b200 import rcps
    from im2im_uq_b200.conv import head_conv_tc, head_conv_tc_hist
    m, packed, hb, g = _head_inputs(B, H, W, seed=B * 1000 + H)
    planes = head_conv_tc(m, packed, hb, 3)                                 # fp32 [B, 3, H, W]
    labels = (planes[:, 1:2] + 0.6 * torch.randn(B, 1, H, W, generator=g).to(DEV)).contiguous()
    labels[0, 0, 0, :4] = float("nan")                                      # NaN labels never miss
    labels[-1, 0, -1, -3:] = planes[-1, 1, -1, -3:]                         # label == prediction: inside every set
    lam = torch.linspace(0.0, lam_hi, L, device=DEV)
    want_counts, want_totals = rcps.miss_counts(planes.view(B, 3, 1, H, W), labels, lam)

    hist = torch.zeros((B, L + 1), dtype=torch.int32, device=DEV)
    planes2 = torch.full_like(planes, float("nan"))
    head_conv_tc_hist(m, packed, hb, labels, lam, hist, out=planes2)
    assert torch.equal(planes2, planes)                                     # the planes it ranked are the planes it stores
    assert int(hist[:, 0].abs().sum()) == 0 and int(hist.sum()) <= B * H * W
    counts = torch.full((B, L), -1, dtype=torch.int32, device=DEV)
    totals = torch.zeros(L, dtype=torch.int64, device=DEV)
    rcps.counts_from_hist(hist, counts, totals)
    assert torch.equal(counts, want_counts) and torch.equal(totals, want_totals)
    assert int(hist.abs().sum()) == 0                                       # self-cleaning: ready for the next batch

    # without the planes at all, twice into the same histogram buffer; the totals keep accumulating
    for rep in (2, 3):
        head_conv_tc_hist(m, packed, hb, labels, lam, hist)
        rcps.counts_from_hist(hist, counts, totals)
        assert torch.equal(counts, want_counts) and torch.equal(totals, rep * want_totals)


def test_head_histogram_refuses_what_it_cannot_rank():
    from im2im_uq_b200 import _lib
    from im2im_uq_b200.conv import head_conv_tc_hist
    m, packed, hb, g = _head_inputs(1, 16, 8, seed=3)
    labels = torch.zeros(1, 1, 16, 8, device=DEV)
    lam = torch.linspace(0, 1, 20000, device=DEV)                           # does not fit in shared memory
    hist = torch.zeros((1, 20001), dtype=torch.int32, device=DEV)
    with pytest.raises(_lib.Im2ImError):
        head_conv_tc_hist(m, packed, hb, labels, lam, hist)
    lib = _lib.load()
    lam = torch.linspace(0, 1, 10, device=DEV)
    hist = torch.zeros((1, 11), dtype=torch.int32, device=DEV)
    rc = lib.im2im_head_conv3x3_tc_hist(m.data_ptr(), packed.data_ptr(), hb.data_ptr(), None, 1, 16, 8, 2, 0, 0, None,
                                        labels.data_ptr(), lam.data_ptr(), 10, hist.data_ptr(), None)
    assert rc == -95                                         # IM2IM_ENOTSUP: two-plane heads are not ranked here


PARAMS = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)


def _model(seed=0):
    from core.models.add_uncertainty import add_uncertainty
    from core.models.trunks.unet import UNet
    torch.manual_seed(seed)
    model = add_uncertainty(UNet(1, 1), PARAMS)
    gen = torch.Generator().manual_seed(seed + 1)
    model.train()
    with torch.no_grad():
        for _ in range(3):
            model(torch.randn(4, 1, 32, 32, generator=gen))
    return model.eval().to(DEV), gen


@pytest.mark.parametrize("shape,fused", [((32, 32), True), ((64, 48), True), ((40, 36), False)])
def test_streaming_calibration_equals_two_stage(shape, fused):
    """calibrate_model batch by batch without the head tensor == collect_outputs + calibrate_from_outputs, bit for bit;
    shapes the tensor-core head cannot tile fall back to per-batch outputs (still no (N, 3, C, H, W) tensor)."""
    from im2im_uq_b200.calibration import calibrate_model as cm
    model, gen = _model()
    n = 29                                                                  # ragged last batch
    x = torch.randn(n, 1, *shape, generator=gen)
    with torch.no_grad():                      # labels scattered around the prediction at ~40 interval half-widths
        o = torch.cat([model(x[i:i + 8].to(DEV)) for i in range(0, n, 8)]).cpu()
    s = 40.0 * torch.randn(n, 1, *shape, generator=gen)   # (an untrained head has upper < prediction on many pixels: zero width)
    y = o[:, 1] + torch.relu(s) * torch.relu(o[:, 2] - o[:, 1]) - torch.relu(-s) * torch.relu(o[:, 1] - o[:, 0])
    ds = torch.utils.data.TensorDataset(x, y)
    cfg = dict(PARAMS, alpha=0.4, delta=0.1, device=DEV, minimum_lambda=0.0, maximum_lambda=60.0, num_lambdas=300,
               rcps_loss="fraction_missed", dataset="synthetic", batch_size=8)
    stats = {}
    with torch.no_grad():
        model, table_s = cm.calibrate_streaming(model, ds, cfg, torch.device(DEV), stats=stats)
        lhat_s = model.lhat.clone()
        assert stats["streaming"] and (stats["head_fused_batches"] == 4) == fused, stats
        model.set_lhat(None)
        model, table_2 = cm.calibrate_model(model, ds, dict(cfg, streaming_calibration=False))
    assert torch.equal(lhat_s, model.lhat) and torch.equal(table_s, table_2)
    assert float(table_s.max()) > 0 and 0 < float(lhat_s) < 60.0             # a non-trivial sweep
    model, table_d = cm.calibrate_model(model, ds, cfg)                      # the default route is the streaming one
    assert torch.equal(table_d, table_2)
    # eval.py:84-126 (dense table at lambdas[j], no early stop) takes the same streaming route
    from core.scripts.eval import get_loss_table
    dense_s = get_loss_table(model, ds, cfg)
    dense_2 = get_loss_table(model, ds, dict(cfg, streaming_calibration=False))
    assert torch.equal(dense_s, dense_2) and float(dense_s[:, 0].min()) > 0


def test_streaming_calibration_descending_grid_and_dataloader_dataset():
    from im2im_uq_b200.calibration import calibrate_model as cm
    model, gen = _model(3)
    n = 12
    x = torch.randn(n, 1, 32, 32, generator=gen)
    y = x + 0.3 * torch.randn(n, 1, 32, 32, generator=gen)

    class Pairs(torch.utils.data.Dataset):                                   # map-style, not a TensorDataset
        def __len__(self):
            return n

        def __getitem__(self, i):
            return x[i], y[i]

    cfg = dict(PARAMS, alpha=0.4, delta=0.1, device=DEV, minimum_lambda=60.0, maximum_lambda=0.0, num_lambdas=120,
               rcps_loss="fraction_missed", dataset="synthetic", batch_size=5)
    with torch.no_grad():
        model, t1 = cm.calibrate_model(model, Pairs(), cfg)
        l1 = model.lhat.clone()
        model, t2 = cm.calibrate_model(model, Pairs(), dict(cfg, streaming_calibration=False))
    assert torch.equal(l1, model.lhat) and torch.equal(t1, t2)


@pytest.mark.parametrize("B,H,W", [(2, 32, 24), (1, 16, 8), (3, 48, 64)])
def test_outconv_folded_into_the_head_equals_the_two_convolutions(B, H, W):
    """head(OutConv(x)) as one tensor-core convolution (im2im_head_conv3x3_tc_folded_f32): equal to the fp32 composition of
    the reference's two modules (unet_parts.py:87-93 -> quantile_layer.py:19-21) up to the bf16 rounding of the operands,
    on the image border as well as inside (OutConv's bias must not leak into the zero padding)."""
    import torch.nn.functional as F
    from im2im_uq_b200.conv import fold_outconv_into_head, head_conv_tc, pack_conv_weight, pad_head_weight
    g = torch.Generator().manual_seed(B * 100 + W)
    y = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16)
    ow, ob = torch.randn(32, 64, 1, 1, generator=g) * 0.1, torch.randn(32, generator=g) * 3.0   # a LARGE OutConv bias
    hw, hb = torch.randn(3, 32, 3, 3, generator=g) * 0.1, torch.randn(3, generator=g)
    ref = F.conv2d(F.conv2d(y.float().permute(0, 3, 1, 2), ow, ob), hw, hb, padding=1)
    wf, bf, tb = fold_outconv_into_head(hw.to(DEV), hb.to(DEV), ow.to(DEV), ob.to(DEV))
    packed = pack_conv_weight(pad_head_weight(wf))
    got = head_conv_tc(y.to(DEV), packed, bf, 3, tap_bias=tb).cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    border = torch.ones(H, W, dtype=torch.bool)
    border[1:-1, 1:-1] = False
    assert err.max().item() <= 1e-2 * scale, (err.max().item(), scale)
    assert err[..., border].max().item() <= 1e-2 * scale
    naive = head_conv_tc(y.to(DEV), packed, bf, 3).cpu()                      # without the border correction: visibly wrong
    assert (naive - ref).abs()[..., border].max().item() > 0.1 * scale
    assert torch.equal(naive[..., ~border], got[..., ~border])              # interior pixels: the very same arithmetic
