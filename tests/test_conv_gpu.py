"""GPU (-m gpu): the tcgen05 implicit-GEMM convolution (C ABI im2im_conv_igemm_bf16) against a plain PyTorch fp32
reference of the same op on the same bf16-rounded operands.

Tolerance: inputs/weights are bf16 (exactly representable in the fp32 reference), accumulation is fp32 in TMEM, so the
only differences are summation order (<= 1e-4 relative to the output scale, checked on the fp32 output) and the final
bf16 rounding of the output (2^-9 relative, checked as 1e-2 of the output scale)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from im2im_uq_b200 import _lib, conv
    DEV = torch.device("cuda:0")


def _ref_and_got(B, H, W, c1, c2, cout, taps, relu, bias, out_dtype, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    k = 3 if taps == 9 else 1
    x1 = torch.randn(B, c1, H, W, device=DEV, generator=g).to(torch.bfloat16)
    x2 = torch.randn(B, c2, H, W, device=DEV, generator=g).to(torch.bfloat16) if c2 else None
    w = (torch.randn(cout, c1 + c2, k, k, device=DEV, generator=g) / ((c1 + c2) * taps) ** 0.5).to(torch.bfloat16)
    b = torch.randn(cout, device=DEV, generator=g) if bias else None
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xin = x1.float() if not c2 else torch.cat([x1.float(), x2.float()], dim=1)  # unet_parts.py:68 order
        ref = F.conv2d(xin, w.float(), b, padding=k // 2)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    if relu:
        ref = ref.relu()
    got = conv.conv_igemm(conv.to_nhwc_bf16(x1), conv.pack_conv_weight(w), b, relu,
                          conv.to_nhwc_bf16(x2) if c2 else None, out_dtype)
    return ref, got.float().permute(0, 3, 1, 2)


@pytest.mark.parametrize("shape", [
    (1, 16, 16, 64, 0, 64, 9),       # one tile
    (2, 40, 40, 128, 0, 128, 9),     # box 8x8x2 tiling of a 40x40 map
    (3, 20, 20, 256, 0, 512, 9),     # batch-padded boxes, two N tiles of 256
    (2, 32, 32, 64, 64, 64, 9),      # concatenated skip + upsampled inputs
    (1, 37, 29, 64, 0, 256, 9),      # ragged map: TMA zero fill on every border
    (2, 16, 8, 64, 0, 32, 1),        # 1x1 OutConv (64 -> 32)
    (1, 320, 320, 64, 0, 64, 9),     # the reference's full-resolution layer
    (5, 20, 20, 512, 512, 512, 9),   # up1 first conv (1024 -> 512), batch not a multiple of the box
    (2, 48, 40, 64, 0, 128, 9),      # halo kernel, resident 64x128 weights (bn = 128)
    (1, 32, 16, 128, 0, 128, 9),     # halo kernel, two N blocks of 64 (weights of one block resident per CTA)
    (3, 16, 8, 64, 64, 64, 9),       # halo kernel with concatenated inputs, single tile per image
    (2, 32, 32, 256, 0, 128, 9),     # halo kernel, CTA pair with STREAMED weights (256 input channels: nothing stays resident)
    (1, 48, 40, 128, 128, 256, 9),   # ... concatenated inputs, two N blocks of 128
    (3, 32, 16, 256, 0, 256, 9),     # ... odd number of tiles
])
def test_conv_matches_fp32_reference(shape):
    B, H, W, c1, c2, cout, taps = shape
    ref, got = _ref_and_got(B, H, W, c1, c2, cout, taps, relu=False, bias=True, out_dtype=torch.float32)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 1e-4 * scale
    ref, got = _ref_and_got(B, H, W, c1, c2, cout, taps, relu=True, bias=True, out_dtype=torch.bfloat16, seed=1)
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert float(got.min()) >= 0.0


def test_conv_argument_validation():
    x = torch.zeros(1, 8, 8, 48, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(64, 9, 48, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(_lib.Im2ImError, match="multiples of 64"):
        conv.conv_igemm(x, w)
    x = torch.zeros(1, 8, 8, 64, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(24, 9, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(_lib.Im2ImError, match="c_out"):
        conv.conv_igemm(x, w)


@pytest.mark.parametrize("B,H,W,c1,c2,cout,halo", [(2, 32, 48, 64, 0, 64, True), (1, 48, 40, 128, 0, 128, True),
                                                   (3, 16, 16, 64, 64, 64, True), (2, 64, 24, 64, 0, 128, True),
                                                   (2, 20, 20, 256, 0, 256, False)])
def test_conv_with_fused_maxpool_equals_conv_then_pool(B, H, W, c1, c2, cout, halo):
    """im2im_conv_igemm_bf16_pool: the convolution's epilogue also writes maxpool2x2 of its output (Down = MaxPool2d(2) ->
    DoubleConv, unet_parts.py:27-36); full-resolution output and pooled tensor bit-identical to conv -> maxpool kernel."""
    g = torch.Generator(device=DEV).manual_seed(11)
    x1 = conv.to_nhwc_bf16(torch.randn(B, c1, H, W, device=DEV, generator=g))
    x2 = conv.to_nhwc_bf16(torch.randn(B, c2, H, W, device=DEV, generator=g)) if c2 else None
    w = conv.pack_conv_weight(torch.randn(cout, c1 + c2, 3, 3, device=DEV, generator=g) / (9 * (c1 + c2)) ** 0.5)
    b = torch.randn(cout, device=DEV, generator=g)
    want = conv.conv_igemm(x1, w, b, True, x2)
    got, pooled = conv.conv_igemm_pool(x1, w, b, True, x2)
    assert torch.equal(got, want)
    assert (pooled is not None) == halo
    if halo:
        ref = torch.empty((B, H // 2, W // 2, cout), dtype=torch.bfloat16, device=DEV)
        _lib.check(_lib.load().im2im_maxpool2x2_bf16(want.data_ptr(), B, H, W, cout, ref.data_ptr(),
                                                     torch.cuda.current_stream(DEV).cuda_stream), "maxpool")
        assert torch.equal(pooled, ref)
        assert float(pooled.float().abs().max()) > 0


@pytest.mark.parametrize("B,H,W,c1,c2,cout", [
    (1, 16, 8, 64, 0, 64),       # one tile: the pair's second tile lies behind the batch (TMA zero fill, epilogue skips it)
    (3, 16, 8, 64, 0, 64),       # odd number of tiles
    (2, 32, 32, 64, 64, 64),     # concatenated inputs: two tensor maps, two channel blocks per tile
    (1, 48, 40, 64, 0, 128),     # N = 128: 64 weight rows per CTA
    (2, 32, 32, 128, 0, 128),    # weights of an N = 128 block fit only when halved (two ring stages)
    (3, 48, 40, 128, 0, 256),    # ... with two such N blocks
    (4, 160, 160, 64, 0, 64),    # more tile pairs than clusters: persistent loop, both TMEM buffers, ring wrap-around
])
def test_halo_conv_cta_pair_equals_single_cta(B, H, W, c1, c2, cout, monkeypatch):
    """The halo convolution runs as clusters of two CTAs (tcgen05 cta_group::2, M = 256; each CTA keeps half of the weight
    rows resident, the leader issues the MMAs for both).  Every output element is accumulated over the same K order as in
    the single-CTA kernel (IM2IM_HALO_PAIR=0), so outputs - fp32, bf16 + ReLU, with the fused max-pool - are bit-identical;
    the fused BatchNorm statistics are fp32 atomics issued in another order (relative 1e-5)."""
    g = torch.Generator(device=DEV).manual_seed(7)
    x1 = conv.to_nhwc_bf16(torch.randn(B, c1, H, W, device=DEV, generator=g))
    x2 = conv.to_nhwc_bf16(torch.randn(B, c2, H, W, device=DEV, generator=g)) if c2 else None
    w = conv.pack_conv_weight(torch.randn(cout, c1 + c2, 3, 3, device=DEV, generator=g) / (9 * (c1 + c2)) ** 0.5)
    b = torch.randn(cout, device=DEV, generator=g)

    def run():
        out = [conv.conv_igemm(x1, w, b, False, x2, torch.float32), conv.conv_igemm(x1, w, b, True, x2)]
        out += list(conv.conv_igemm_pool(x1, w, b, True, x2))
        sums = torch.zeros(2 * cout, device=DEV)
        z, fused = conv.conv_igemm_stats(x1, w, 1, sums, x2=x2)
        torch.cuda.synchronize()
        return out + [z], sums, fused

    monkeypatch.setenv("IM2IM_HALO_PAIR", "0")
    single, sums_single, fused_single = run()
    monkeypatch.setenv("IM2IM_HALO_PAIR", "1")
    before = _lib.launch_count()
    pair, sums_pair, fused_pair = run()
    assert _lib.launch_count() - before == 4
    for a, p in zip(single, pair):
        assert (a is None) == (p is None)
        if a is not None:
            assert torch.equal(a, p)
    assert fused_single == fused_pair
    if fused_pair:
        assert torch.allclose(sums_single, sums_pair, rtol=1e-5, atol=1e-5 * float(sums_single.abs().max()))
    # and against the fp32 reference of the op (the parity yardstick of this file)
    xin = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], dim=3)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(xin.permute(0, 3, 1, 2), w.float().view(cout, 3, 3, c1 + c2).permute(0, 3, 1, 2), b, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert (pair[0].permute(0, 3, 1, 2) - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("B,H,W,cin,cout", [
    (1, 16, 16, 64, 64),       # halo weight gradient, one pixel tile per image, N = 64
    (2, 40, 40, 128, 64),      # split-K kernel (40x40 is not an 8x16-tile shape)
    (3, 20, 20, 256, 512),     # split-K kernel, deep layer
    (2, 48, 24, 128, 192),     # halo, several channel / c_out blocks, c_out not a multiple of 128
    (2, 64, 64, 128, 128),     # halo, N = 128: taps dealt to two kinds of CTA
    (3, 32, 40, 128, 256),     # ... two c_out blocks of 128
    (1, 48, 40, 256, 128),     # ... four channel blocks
])
def test_conv_weight_gradient_matches_autograd(B, H, W, cin, cout, monkeypatch):
    """im2im_conv_wgrad_bf16 (dW of the 3x3 convolutions of unet_parts.py:16-21) against torch autograd in fp32 on the same
    bf16-rounded operands; the N = 128 form of the halo kernel against its N = 64 form (fp32 atomics in another order)."""
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(B, cin, H, W, device=DEV, generator=g).to(torch.bfloat16)
    dz = torch.randn(B, cout, H, W, device=DEV, generator=g).to(torch.bfloat16)
    wf = torch.zeros(cout, cin, 3, 3, device=DEV, requires_grad=True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        F.conv2d(x.float(), wf, None, padding=1).backward(dz.float())
    finally:
        torch.backends.cudnn.allow_tf32 = old
    ref = wf.grad.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    scale = ref.abs().max().item()
    xn, dzn = conv.to_nhwc_bf16(x), conv.to_nhwc_bf16(dz)
    got = conv.conv_wgrad(xn, dzn, 9)
    assert (got - ref).abs().max().item() <= 5e-4 * scale
    monkeypatch.setenv("IM2IM_WGRAD_HALO_WIDE", "0")
    narrow = conv.conv_wgrad(xn, dzn, 9)
    assert (narrow - ref).abs().max().item() <= 5e-4 * scale
    assert (narrow - got).abs().max().item() <= 1e-4 * scale
    # accumulates into `out`
    again = conv.conv_wgrad(xn, dzn, 9, out=got.clone())
    assert (again - 2 * ref).abs().max().item() <= 1e-3 * scale
