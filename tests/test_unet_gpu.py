"""GPU (-m gpu): native sm_100a inference forward (tcgen05 convolutions, bf16 activations, fp32 accumulation) of
UNet + quantile head against (1) the reference's own fp32 output (tests/golden/unet_forward_kat.npz, produced by the
unmodified reference) and (2) the fp32 PyTorch forward of the same module.

Tolerance: activations and folded weights are rounded to bf16 (2^-9 relative) at each of the 23 layers while the
reference is fp32 end to end, so outputs agree to ~1e-2 of the output scale; asserted: max error <= 4e-2 * scale and
relative L2 error <= 2e-2."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PARAMS = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
MAX_TOL, L2_TOL = 4e-2, 2e-2


def _seeded_model(c_in=1, c_out=1, warm=True):
    from core.models.add_uncertainty import add_uncertainty
    from core.models.trunks.unet import UNet
    torch.manual_seed(0)
    model = add_uncertainty(UNet(c_in, c_out), PARAMS)
    gen = torch.Generator().manual_seed(1)
    if warm:  # move the BatchNorm running statistics away from (0, 1) exactly as make_golden.py does
        model.train()
        with torch.no_grad():
            for _ in range(3):
                model(torch.randn(4, c_in, 32, 32, generator=gen))
    model.eval()
    return model, gen


def _errors(got, want):
    scale = want.abs().max().item()
    return (got - want).abs().max().item() / scale, ((got - want).norm() / want.norm()).item()


def test_native_forward_matches_reference_output():
    g = np.load(os.path.join(GOLDEN, "unet_forward_kat.npz"))
    model, gen = _seeded_model()
    sd = model.state_dict()
    assert list(sd.keys()) == json.loads(str(g["keys"]))                     # checkpoint-compatible names
    checksum = float(sum(v.double().abs().sum() for v in sd.values() if v is not None and v.dtype.is_floating_point))
    assert checksum == float(g["state_checksum"])                            # same seeded state as the reference run
    x = torch.randn(2, 1, 48, 32, generator=gen)
    assert np.array_equal(x.numpy(), g["x"])
    model = model.to("cuda:0")
    with torch.no_grad():
        y = model(x.to("cuda:0"))
    assert "_native_engine" in model.__dict__                                # the CUDA path ran, not the module graph
    assert tuple(y.shape) == tuple(g["y"].shape) and y.dtype == torch.float32
    mx, l2 = _errors(y.cpu(), torch.from_numpy(g["y"]))
    assert mx <= MAX_TOL and l2 <= L2_TOL, (mx, l2)


@pytest.mark.parametrize("shape,c_in,c_out", [((3, 64, 64), 1, 1), ((2, 50, 38), 1, 1), ((1, 320, 320), 1, 1),
                                               ((2, 32, 48), 2, 2)])
def test_native_forward_matches_fp32_module(shape, c_in, c_out):
    model, gen = _seeded_model(c_in, c_out)
    model = model.to("cuda:0")
    b, h, w = shape
    x = torch.randn(b, c_in, h, w, generator=gen).to("cuda:0")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            model.use_native_inference = False
            ref = model(x)
            model.use_native_inference = True
            got = model(x)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert tuple(got.shape) == (b, 3, c_out, h, w)
    mx, l2 = _errors(got, ref)
    assert mx <= MAX_TOL and l2 <= L2_TOL, (mx, l2)


def test_engine_follows_weight_updates_and_training_mode():
    model, gen = _seeded_model()
    model = model.to("cuda:0")
    x = torch.randn(2, 1, 32, 32, generator=gen).to("cuda:0")
    with torch.no_grad():
        y0 = model(x)
        for p in model.parameters():
            p.mul_(1.01)
        y1 = model(x)
        model.use_native_inference = False
        ref1 = model(x)
        model.use_native_inference = True
    assert not torch.equal(y0, y1)
    mx, l2 = _errors(y1, ref1)
    assert mx <= MAX_TOL and l2 <= L2_TOL
    model.train()
    out = model(x)                     # training forward goes through autograd-visible modules
    assert out.requires_grad


def test_calibrate_model_end_to_end_with_native_unet():
    """router.py-style flow: model -> calibrate_model(model, dataset, config) with the UNet forward and the RCPS sweep
    both native; lhat/table must equal the oracle's sweep over the very outputs the native forward produced."""
    from core.calibration.calibrate_model import calibrate_model
    from oracle import rcps_oracle as orc
    model, gen = _seeded_model()
    n = 24
    x = torch.randn(n, 1, 32, 32, generator=gen)
    y = x + 0.3 * torch.randn(n, 1, 32, 32, generator=gen)
    ds = torch.utils.data.TensorDataset(x, y)
    cfg = dict(PARAMS, alpha=0.4, delta=0.1, device="cuda:0", minimum_lambda=0.0, maximum_lambda=60.0, num_lambdas=300,
               rcps_loss="fraction_missed", dataset="synthetic", batch_size=8)
    model, table = calibrate_model(model, ds, cfg)
    with torch.no_grad():
        outs = torch.cat([model(x[i:i + 8].to("cuda:0")) for i in range(0, n, 8)]).cpu()
    lhat, stop, ref_table = orc.calibrate_sweep(outs.numpy(), y.numpy(), 0.0, 60.0, 300, 0.4, 0.1)
    assert torch.equal(model.lhat, lhat) and torch.equal(table, ref_table)


# ------------------------------------------------------------------------------------------- reference precision (tf32)
TF32_L2_TOL, TF32_MAX_TOL = 2e-3, 6e-3   # kind::tf32 (10-bit mantissa operands, fp32 accumulation) over 23 layers vs fp32


@pytest.mark.parametrize("c1,c2,cout,taps,shape", [(64, 0, 64, 9, (2, 32, 48)), (128, 64, 128, 9, (1, 40, 40)),
                                                   (32, 0, 32, 1, (3, 20, 20)), (512, 512, 256, 9, (2, 16, 16)),
                                                   (64, 0, 32, 1, (2, 50, 38)),
                                                   # halo kernel as CTA pairs (8x16-tile shapes): resident weights with the
                                                   # doubled budget, N = 128, odd tile count, streamed weights, concatenation
                                                   (128, 0, 64, 9, (1, 48, 40)), (64, 0, 128, 9, (3, 16, 8)),
                                                   (128, 0, 128, 9, (2, 32, 32)), (64, 64, 256, 9, (1, 32, 16))])
def test_conv_igemm_tf32_vs_torch_fp32(c1, c2, cout, taps, shape):
    """The kind::tf32 convolution on TF32-representable inputs is an exact-product / fp32-accumulate GEMM: it must agree
    with torch's fp32 convolution (TF32 off) to accumulation-order rounding."""
    from im2im_uq_b200.conv import conv_igemm_tf32, pack_conv_weight_tf32, round_to_tf32
    b, h, w = shape
    g = torch.Generator(device="cuda:0").manual_seed(3)
    k = 3 if taps == 9 else 1
    x1 = round_to_tf32(torch.randn(b, c1, h, w, device="cuda:0", generator=g))
    x2 = round_to_tf32(torch.randn(b, c2, h, w, device="cuda:0", generator=g)) if c2 else None
    wt = round_to_tf32(torch.randn(cout, c1 + c2, k, k, device="cuda:0", generator=g) / ((c1 + c2) * taps) ** 0.5)
    bias = torch.randn(cout, device="cuda:0", generator=g)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xin = x1 if x2 is None else torch.cat([x1, x2], dim=1)
        ref = torch.relu(torch.nn.functional.conv2d(xin, wt, bias, padding=k // 2))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    got = conv_igemm_tf32(nhwc(x1), pack_conv_weight_tf32(wt), bias, relu=True, x2=nhwc(x2) if x2 is not None else None)
    got = got.permute(0, 3, 1, 2)
    # outputs are rounded onto the TF32 grid on store (half an ulp = 2^-11 = 4.9e-4 relative); the rest is fp32 accumulation order
    assert ((got - ref).abs() <= 5.2e-4 * ref.abs() + 1e-4).all(), float((got - ref).abs().max())


def test_tf32_forward_matches_reference_output():
    """Reference-precision mode against the reference's own fp32 output (fixture from the unmodified reference)."""
    g = np.load(os.path.join(GOLDEN, "unet_forward_kat.npz"))
    model, gen = _seeded_model()
    x = torch.randn(2, 1, 48, 32, generator=gen)
    assert np.array_equal(x.numpy(), g["x"])
    model = model.to("cuda:0")
    model.native_precision = "tf32"
    from im2im_uq_b200 import _lib
    before = _lib.launch_count()
    with torch.no_grad():
        y = model(x.to("cuda:0"))
    assert _lib.launch_count() - before >= 25 and model.__dict__["_native_engine"].precision == "tf32"
    mx, l2 = _errors(y.cpu(), torch.from_numpy(g["y"]))
    assert mx <= TF32_MAX_TOL and l2 <= TF32_L2_TOL, (mx, l2)
    # and it is an order of magnitude closer to fp32 than the bf16 mode
    model.native_precision = "bf16"
    with torch.no_grad():
        y16 = model(x.to("cuda:0"))
    assert model.__dict__["_native_engine"].precision == "bf16"
    mx16, l216 = _errors(y16.cpu(), torch.from_numpy(g["y"]))
    assert l2 < 0.25 * l216, (l2, l216)


@pytest.mark.parametrize("shape,c_in,c_out", [((3, 64, 64), 1, 1), ((2, 50, 38), 1, 1), ((1, 320, 320), 1, 1),
                                               ((2, 32, 48), 2, 2)])
def test_tf32_forward_matches_fp32_module(shape, c_in, c_out):
    model, gen = _seeded_model(c_in, c_out)
    model = model.to("cuda:0")
    model.native_precision = "tf32"
    b, h, w = shape
    x = torch.randn(b, c_in, h, w, generator=gen).to("cuda:0")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            model.use_native_inference = False
            ref = model(x)
            model.use_native_inference = True
            got = model(x)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert tuple(got.shape) == (b, 3, c_out, h, w)
    mx, l2 = _errors(got, ref)
    assert mx <= TF32_MAX_TOL and l2 <= TF32_L2_TOL, (mx, l2)
