"""GPU (-m gpu): each training-side kernel against the PyTorch fp32 reference of the same op (autograd where relevant).

Tolerances: tensors that are stored in bf16 carry 2^-9 relative rounding (checked as <= 1e-2 of the tensor's scale);
fp32 reductions are checked to 1e-4..1e-3 relative."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from im2im_uq_b200 import _lib
    from im2im_uq_b200.models import unet_train
    DEV = torch.device("cuda:0")
    LIB = _lib.load()


def _st():
    return torch.cuda.current_stream(DEV).cuda_stream


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _close(got, want, tol):
    scale = want.abs().max().item() + 1e-30
    err = (got - want).abs().max().item()
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("shape", [(3, 64, 20, 12), (2, 128, 16, 16), (1, 512, 8, 8)])
def test_batchnorm_relu_forward_backward(shape):
    B, C, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(0)
    z = (torch.randn(B, C, H, W, device=DEV, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    dy = torch.randn(B, C, H, W, device=DEV, generator=g).to(torch.bfloat16)
    bn = torch.nn.BatchNorm2d(C).to(DEV).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.uniform_(-0.5, 0.5, generator=g)
    conv = torch.nn.Conv2d(C, C, 3, padding=1).to(DEV)
    bn_ref = torch.nn.BatchNorm2d(C).to(DEV).train()
    bn_ref.load_state_dict(bn.state_dict())
    zf = z.float().requires_grad_(True)
    y_ref = F.relu(bn_ref(zf + conv.bias.detach()[None, :, None, None]))
    y_ref.backward(dy.float())
    eng = unet_train.UNetTrainEngine.__new__(unet_train.UNetTrainEngine)
    eng.lib = LIB
    layer = unet_train._ConvBN(conv, bn)
    saved, grads = {}, {}
    y = eng._bn_relu(_nhwc(z), layer, saved)
    dz = eng._bn_relu_bwd(_nhwc(dy), layer, saved, grads)
    _close(_nchw(y), y_ref.detach(), 1e-2)
    _close(_nchw(dz), zf.grad, 2e-2)
    _close(grads[bn.weight], bn_ref.weight.grad, 2e-2)
    _close(grads[bn.bias], bn_ref.bias.grad, 1e-2)
    _close(bn.running_mean, bn_ref.running_mean, 1e-3)
    _close(bn.running_var, bn_ref.running_var, 1e-3)
    assert int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (1, 128, 10, 14), (2, 64, 9, 7)])
def test_maxpool_backward(shape):
    B, C, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(B, C, H, W, device=DEV, generator=g).to(torch.bfloat16)
    x[:, :, :4, :4] = 0.25  # ties: gradient must go to the first maximal element like ATen
    dy = torch.randn(B, C, H // 2, W // 2, device=DEV, generator=g).to(torch.bfloat16)
    xf = x.float().requires_grad_(True)
    F.max_pool2d(xf, 2).backward(dy.float())
    prev = torch.randn(B, H, W, C, device=DEV, generator=g).to(torch.bfloat16)
    for accumulate in (0, 1):
        dx = prev.clone() if accumulate else torch.full((B, H, W, C), 7.0, device=DEV, dtype=torch.bfloat16)
        _lib.check(LIB.im2im_maxpool2x2_bwd_bf16(_nhwc(x).data_ptr(), _nhwc(dy).data_ptr(), B, H, W, C, accumulate,
                                                 dx.data_ptr(), _st()), "maxpool_bwd")
        want = xf.grad + (_nchw(prev) if accumulate else 0)
        _close(_nchw(dx), want, 1e-2)


@pytest.mark.parametrize("shape", [(2, 64, 8, 8, 16, 16), (1, 64, 5, 7, 11, 15), (2, 128, 20, 20, 40, 40)])
def test_upsample_forward_backward(shape):
    B, C, h, w, Ho, Wo = shape
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(B, C, h, w, device=DEV, generator=g).to(torch.bfloat16)
    du = torch.randn(B, C, Ho, Wo, device=DEV, generator=g).to(torch.bfloat16)
    xf = x.float().requires_grad_(True)
    up = F.interpolate(xf, scale_factor=2, mode="bilinear", align_corners=True)
    dyy, dxx = Ho - up.shape[2], Wo - up.shape[3]
    up = F.pad(up, [dxx // 2, dxx - dxx // 2, dyy // 2, dyy - dyy // 2])  # unet_parts.py:63-64
    up.backward(du.float())
    y = torch.empty((B, Ho, Wo, C), device=DEV, dtype=torch.bfloat16)
    _lib.check(LIB.im2im_upsample2x_bilinear_bf16(_nhwc(x).data_ptr(), B, h, w, C, Ho, Wo, y.data_ptr(), _st()), "up")
    _close(_nchw(y), up.detach(), 1e-2)
    dx = torch.empty((B, h, w, C), device=DEV, dtype=torch.bfloat16)
    _lib.check(LIB.im2im_upsample2x_bilinear_bwd_bf16(_nhwc(du).data_ptr(), B, h, w, C, Ho, Wo, dx.data_ptr(), _st()),
               "up_bwd")
    _close(_nchw(dx), xf.grad, 1e-2)


@pytest.mark.parametrize("c_out", [1, 2])
def test_head_forward_backward(c_out):
    B, H, W, c_mid, n_out = 2, 12, 10, 32, 3 * c_out
    g = torch.Generator(device=DEV).manual_seed(3)
    m = torch.randn(B, c_mid, H, W, device=DEV, generator=g).to(torch.bfloat16)
    w = torch.randn(n_out, c_mid, 3, 3, device=DEV, generator=g) * 0.1
    b = torch.randn(n_out, device=DEV, generator=g)
    dout = torch.randn(B, n_out, H, W, device=DEV, generator=g)
    mf, wf, bf = m.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.conv2d(mf, wf, bf, padding=1).backward(dout)
    m_pad = torch.zeros((B, H, W, 64), device=DEV, dtype=torch.bfloat16)
    m_pad[..., :c_mid] = _nhwc(m)
    dm = torch.full((B, H, W, 64), 3.0, device=DEV, dtype=torch.bfloat16)
    dw = torch.zeros_like(w); db = torch.zeros_like(b)
    _lib.check(LIB.im2im_head_bwd(dout.contiguous().data_ptr(), m_pad.data_ptr(), w.contiguous().data_ptr(), B, H, W,
                                  c_mid, 64, n_out, dm.data_ptr(), dw.data_ptr(), db.data_ptr(), _st()), "head_bwd")
    _close(_nchw(dm[..., :c_mid]), mf.grad, 1e-2)
    assert float(dm[..., c_mid:].abs().max()) == 0.0
    _close(dw, wf.grad, 1e-4)
    _close(db, bf.grad, 1e-4)


def test_first_conv_forward_and_wgrad():
    B, c_in, H, W, c_out = 3, 1, 20, 18, 64
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(B, c_in, H, W, device=DEV, generator=g)
    w = torch.randn(c_out, c_in, 3, 3, device=DEV, generator=g)
    dz = torch.randn(B, c_out, H, W, device=DEV, generator=g).to(torch.bfloat16)
    wf = w.clone().requires_grad_(True)
    z_ref = F.conv2d(x, wf, None, padding=1)
    z_ref.backward(dz.float())
    z = torch.empty((B, H, W, c_out), device=DEV, dtype=torch.bfloat16)
    _lib.check(LIB.im2im_conv_first_bf16(x.data_ptr(), w.data_ptr(), None, B, c_in, H, W, c_out, 0, z.data_ptr(),
                                         _st()), "conv_first")
    _close(_nchw(z), z_ref.detach(), 1e-2)
    dw = torch.zeros_like(w)
    _lib.check(LIB.im2im_conv_first_wgrad(x.data_ptr(), _nhwc(dz).data_ptr(), B, c_in, H, W, c_out, dw.data_ptr(),
                                          _st()), "conv_first_wgrad")
    _close(dw, wf.grad, 1e-4)


def test_quantile_loss_matches_reference_kats():
    g = np.load(os.path.join(GOLDEN, "quantile_loss_kats.npz"))
    for name in "abc":
        pred = torch.from_numpy(g[f"{name}_pred"]).to(DEV).requires_grad_(True)
        target = torch.from_numpy(g[f"{name}_target"]).to(DEV)
        params = json.loads(str(g[f"{name}_params"]))
        loss = unet_train.native_quantile_loss(pred, target, params)
        loss.backward()
        np.testing.assert_allclose(loss.item(), float(g[f"{name}_loss"]), rtol=1e-6)
        np.testing.assert_allclose(pred.grad.cpu().numpy(), g[f"{name}_grad"], rtol=1e-6, atol=1e-10)


def test_fused_adam_matches_torch_adam():
    g = torch.Generator(device=DEV).manual_seed(5)
    ps = [torch.nn.Parameter(torch.randn(s, device=DEV, generator=g)) for s in [(64, 3, 3), (17,), (5, 5)]]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref = torch.optim.Adam(qs, lr=1e-3)
    ours = unet_train.FusedAdam(ps, lr=1e-3)
    for step in range(4):
        for p, q in zip(ps, qs):
            gr = torch.randn(p.shape, device=DEV, generator=g) * (10.0 ** (step - 2))
            p.grad.copy_(gr); q.grad = gr.clone()
        ours.step(); ref.step()
        for p, q in zip(ps, qs):
            np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().cpu().numpy(), rtol=2e-6, atol=1e-8)


# ------------------------------------------------------------------------------------------- statistics fused into conv epilogues
@pytest.mark.parametrize("c1,c2,cout,shape", [(64, 0, 64, (2, 32, 48)), (64, 64, 64, (3, 16, 16)), (64, 0, 128, (2, 32, 16)),
                                              (128, 0, 128, (1, 48, 40)), (128, 0, 256, (2, 16, 8)),
                                              # N = 128 tiles with the per-tile shuffle reduction; 256 input channels:
                                              # CTA pair with streamed weights
                                              (256, 0, 128, (2, 32, 16)), (128, 128, 256, (3, 16, 8)), (256, 0, 256, (1, 48, 40))])
def test_conv_epilogue_batch_statistics(c1, c2, cout, shape):
    """im2im_conv_igemm_bf16_stats mode 1: output identical to the plain convolution; sums = channel sums / sums of squares
    of the stored bf16 output (what im2im_channel_stats_bf16 would compute in a second pass)."""
    from im2im_uq_b200.conv import conv_igemm, conv_igemm_stats, pack_conv_weight
    B, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(4)
    x1 = _nhwc(torch.randn(B, c1, H, W, device=DEV, generator=g))
    x2 = _nhwc(torch.randn(B, c2, H, W, device=DEV, generator=g)) if c2 else None
    w = pack_conv_weight(torch.randn(cout, c1 + c2, 3, 3, device=DEV, generator=g) / (9 * (c1 + c2)) ** 0.5)
    sums = torch.zeros(2 * cout, device=DEV)
    z, fused = conv_igemm_stats(x1, w, 1, sums, x2=x2)
    assert fused                                           # every shape here is a halo-kernel layer
    assert torch.equal(z, conv_igemm(x1, w, x2=x2))
    zf = z.float().reshape(-1, cout)
    _close(sums[:cout], zf.sum(0), 2e-5)
    _close(sums[cout:], (zf * zf).sum(0), 2e-5)
    # accumulates: a second call doubles the sums
    conv_igemm_stats(x1, w, 1, sums, x2=x2)
    _close(sums[:cout], 2 * zf.sum(0), 2e-5)


@pytest.mark.parametrize("c1,c2,cout,shape", [(512, 0, 512, (2, 20, 20)), (256, 0, 512, (3, 40, 40)), (512, 512, 512, (1, 40, 40)),
                                              (256, 0, 128, (2, 20, 12)), (256, 256, 256, (1, 10, 10)), (512, 0, 256, (5, 5, 7))])
def test_persistent_conv_epilogue_batch_statistics(c1, c2, cout, shape):
    """The deep layers (not 8x16-tile shapes / weights too large for the halo kernel) run on the persistent kernel; with
    >= 256 input channels its epilogue accumulates the same statistics (recursive-halving column sums + fp32 atomics),
    ragged tiles included (rows outside the image contribute nothing)."""
    from im2im_uq_b200.conv import conv_igemm, conv_igemm_stats, pack_conv_weight
    B, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(5)
    x1 = _nhwc(torch.randn(B, c1, H, W, device=DEV, generator=g))
    x2 = _nhwc(torch.randn(B, c2, H, W, device=DEV, generator=g)) if c2 else None
    w = pack_conv_weight(torch.randn(cout, c1 + c2, 3, 3, device=DEV, generator=g) / (9 * (c1 + c2)) ** 0.5)
    sums = torch.zeros(2 * cout, device=DEV)
    z, fused = conv_igemm_stats(x1, w, 1, sums, x2=x2)
    assert fused
    assert torch.equal(z, conv_igemm(x1, w, x2=x2))
    zf = z.float().reshape(-1, cout)
    _close(sums[:cout], zf.sum(0), 2e-5)
    _close(sums[cout:], (zf * zf).sum(0), 2e-5)
    conv_igemm_stats(x1, w, 1, sums, x2=x2)                                # accumulates
    _close(sums[:cout], 2 * zf.sum(0), 2e-5)


def test_conv_epilogue_statistics_fall_back_for_shallow_layers_off_the_halo_kernel():
    from im2im_uq_b200.conv import conv_igemm, conv_igemm_stats, pack_conv_weight
    g = torch.Generator(device=DEV).manual_seed(5)
    x = _nhwc(torch.randn(2, 128, 20, 20, device=DEV, generator=g))     # 20x20: not an 8x16-tile shape, K loop too short
    w = pack_conv_weight(torch.randn(256, 128, 3, 3, device=DEV, generator=g) / 34)
    sums = torch.zeros(512, device=DEV)
    z, fused = conv_igemm_stats(x, w, 1, sums)
    assert not fused and float(sums.abs().max()) == 0.0 and torch.equal(z, conv_igemm(x, w))


@pytest.mark.parametrize("cin,cout,shape", [(64, 64, (2, 32, 48)), (128, 64, (2, 32, 16)), (128, 128, (1, 48, 40))])
def test_dgrad_epilogue_batchnorm_backward_sums(cin, cout, shape):
    """stat_mode 2: the data-gradient convolution stores g = dy * relu_mask and the two BatchNorm-backward sums; followed by
    the apply pass it must equal the unfused conv -> im2im_bn_relu_bwd_bf16 sequence."""
    from im2im_uq_b200.conv import conv_igemm, conv_igemm_stats, pack_conv_weight
    B, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(6)
    dz_in = _nhwc(torch.randn(B, cin, H, W, device=DEV, generator=g))
    w = pack_conv_weight(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (9 * cin) ** 0.5)
    z = _nhwc(torch.randn(B, cout, H, W, device=DEV, generator=g) * 1.3 + 0.2)
    gamma = torch.rand(cout, device=DEV, generator=g) + 0.5
    beta = torch.rand(cout, device=DEV, generator=g) - 0.5
    zf = z.float().reshape(-1, cout)
    mean = zf.mean(0).contiguous()
    rstd = (1.0 / torch.sqrt(zf.var(0, unbiased=False) + 1e-5)).contiguous()
    n_pix = B * H * W
    # unfused reference path
    dy = conv_igemm(dz_in, w)
    sums_ref = torch.empty(2 * cout, device=DEV)
    dz_ref = torch.empty_like(z)
    _lib.check(LIB.im2im_bn_relu_bwd_bf16(dy.data_ptr(), z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                                          rstd.data_ptr(), n_pix, cout, sums_ref.data_ptr(), dz_ref.data_ptr(), _st()))
    # fused path
    sums = torch.zeros(2 * cout, device=DEV)
    gm, fused = conv_igemm_stats(dz_in, w, 2, sums, bn=(z, gamma, beta, mean, rstd))
    assert fused
    sc = gamma * rstd
    mask = (zf * sc + (beta - mean * sc)) > 0
    assert torch.equal(gm.reshape(-1, cout), torch.where(mask, dy.reshape(-1, cout), torch.zeros_like(dy.reshape(-1, cout))))
    _close(sums, sums_ref, 5e-5)
    dz = torch.empty_like(z)
    _lib.check(LIB.im2im_bn_relu_bwd_apply_bf16(gm.data_ptr(), z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                                                rstd.data_ptr(), sums.data_ptr(), n_pix, cout, 1, dz.data_ptr(), _st()))
    _close(dz.float(), dz_ref.float(), 1e-2)


# ------------------------------------------------------------------------------------------- pool fused into BatchNorm (skip layers)
@pytest.mark.parametrize("shape", [(3, 64, 20, 12), (2, 128, 16, 16), (1, 512, 8, 4), (2, 256, 6, 10)])
def test_bn_relu_pool_forward_and_backward_equal_the_separate_kernels(shape):
    """im2im_bn_apply_relu_pool_bf16 / im2im_bn_relu_pool_bwd_bf16 against the sequences they replace:
    bn_apply_relu -> maxpool2x2 forward, maxpool2x2_bwd(accumulate) -> bn_relu_bwd backward.  Same arithmetic per element,
    so y / pooled / dz are bit-identical; the per-channel sums differ only in summation order."""
    B, C, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(8)
    z = _nhwc(torch.randn(B, C, H, W, device=DEV, generator=g) * 1.5 + 0.3)
    flat = z.view(-1)
    n5 = flat.numel() // 5 * 5
    flat[0:n5:5] = flat[1:n5:5].clone()                                    # ties inside pooling windows
    scale = torch.rand(C, device=DEV, generator=g) + 0.5
    shift = torch.rand(C, device=DEV, generator=g) - 0.7
    n_pix = B * H * W
    y_ref = torch.empty_like(z)
    _lib.check(LIB.im2im_bn_apply_relu_bf16(z.data_ptr(), scale.data_ptr(), shift.data_ptr(), n_pix, C, y_ref.data_ptr(), _st()))
    p_ref = torch.empty((B, H // 2, W // 2, C), dtype=torch.bfloat16, device=DEV)
    _lib.check(LIB.im2im_maxpool2x2_bf16(y_ref.data_ptr(), B, H, W, C, p_ref.data_ptr(), _st()))
    y, p = torch.empty_like(z), torch.empty_like(p_ref)
    _lib.check(LIB.im2im_bn_apply_relu_pool_bf16(z.data_ptr(), scale.data_ptr(), shift.data_ptr(), B, H, W, C, y.data_ptr(),
                                                 p.data_ptr(), _st()))
    assert torch.equal(y, y_ref) and torch.equal(p, p_ref)
    # backward: scale/shift expressed through (gamma, beta, mean, rstd) exactly as the kernels recompute them
    rstd = torch.rand(C, device=DEV, generator=g) + 0.5
    gamma = scale / rstd
    mean = torch.randn(C, device=DEV, generator=g) * 0.1
    beta = shift + mean * (gamma * rstd)
    y2 = torch.empty_like(z)
    sc2 = (gamma * rstd).contiguous()
    sh2 = (beta - mean * sc2).contiguous()
    _lib.check(LIB.im2im_bn_apply_relu_bf16(z.data_ptr(), sc2.data_ptr(), sh2.data_ptr(), n_pix, C, y2.data_ptr(), _st()))
    d_skip = _nhwc(torch.randn(B, C, H, W, device=DEV, generator=g))
    d_p = torch.randn(B, H // 2, W // 2, C, device=DEV, generator=g).to(torch.bfloat16)
    dy = d_skip.clone()
    _lib.check(LIB.im2im_maxpool2x2_bwd_bf16(y2.data_ptr(), d_p.data_ptr(), B, H, W, C, 1, dy.data_ptr(), _st()))
    sums_ref = torch.empty(2 * C, device=DEV)
    dz_ref = torch.empty_like(z)
    _lib.check(LIB.im2im_bn_relu_bwd_bf16(dy.data_ptr(), z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                                          rstd.data_ptr(), n_pix, C, sums_ref.data_ptr(), dz_ref.data_ptr(), _st()))
    sums = torch.empty(2 * C, device=DEV)
    dz = torch.empty_like(z)
    _lib.check(LIB.im2im_bn_relu_pool_bwd_bf16(d_skip.data_ptr(), d_p.data_ptr(), z.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                               mean.data_ptr(), rstd.data_ptr(), B, H, W, C, sums.data_ptr(), dz.data_ptr(), _st()))
    _close(sums, sums_ref, 2e-5)
    _close(dz.float(), dz_ref.float(), 4e-3)          # one bf16 ulp where the sums' last bits move a rounding
    assert (dz != dz_ref).float().mean().item() < 0.02


def test_bn_relu_pool_rejects_odd_sizes():
    z = torch.zeros(1, 5, 4, 64, dtype=torch.bfloat16, device=DEV)
    s = torch.zeros(64, device=DEV)
    rc = LIB.im2im_bn_apply_relu_pool_bf16(z.data_ptr(), s.data_ptr(), s.data_ptr(), 1, 5, 4, 64, z.data_ptr(), z.data_ptr(), _st())
    assert rc == -95          # IM2IM_ENOTSUP: the engine then runs the separate kernels
