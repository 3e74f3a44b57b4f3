"""Tensor-level wrappers over the tcgen05 convolution entry points of the C ABI (include/im2im_uq.h).

Activations are NHWC bf16 CUDA tensors; weights are packed once to [c_out, taps, c_in] bf16 (K-major rows).
"""
from typing import Optional

import torch

from . import _lib


def pack_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """torch Conv2d weight [c_out, c_in, kh, kw] -> bf16 [c_out, kh*kw, c_in] (tap-major K), contiguous."""
    c_out, c_in, kh, kw = weight.shape
    return weight.detach().permute(0, 2, 3, 1).reshape(c_out, kh * kw, c_in).contiguous().to(torch.bfloat16)


def to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """NCHW float -> NHWC bf16 contiguous."""
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def conv_igemm(x1: torch.Tensor, weight_packed: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
               x2: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16) -> torch.Tensor:
    """3x3 (pad 1) or 1x1 convolution on tcgen05: x1 (and optionally x2, concatenated after x1 on channels) NHWC bf16."""
    lib = _lib.load()
    assert x1.is_cuda and x1.dtype == torch.bfloat16 and x1.is_contiguous() and x1.dim() == 4
    B, H, W, c1 = x1.shape
    c2 = 0
    if x2 is not None:
        assert x2.is_cuda and x2.dtype == torch.bfloat16 and x2.is_contiguous() and tuple(x2.shape[:3]) == (B, H, W)
        c2 = x2.shape[3]
    c_out, taps, c_in = weight_packed.shape
    assert weight_packed.dtype == torch.bfloat16 and weight_packed.is_contiguous() and c_in == c1 + c2
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == c_out and bias.is_cuda
    out = torch.empty((B, H, W, c_out), dtype=out_dtype, device=x1.device)
    with torch.cuda.device(x1.device):
        rc = lib.im2im_conv_igemm_bf16(x1.data_ptr(), c1, x2.data_ptr() if x2 is not None else None, c2,
                                       weight_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
                                       B, H, W, c_out, taps, 1 if relu else 0,
                                       out.data_ptr() if out_dtype == torch.bfloat16 else None,
                                       out.data_ptr() if out_dtype == torch.float32 else None,
                                       torch.cuda.current_stream(x1.device).cuda_stream)
    _lib.check(rc, "im2im_conv_igemm_bf16")
    return out


def conv_igemm_stats(x1: torch.Tensor, weight_packed: torch.Tensor, stat_mode: int, sums: torch.Tensor,
                     x2: Optional[torch.Tensor] = None, bn=None):
    """conv_igemm (no bias / ReLU, bf16 out) with per-channel statistics fused into the epilogue
    (im2im_conv_igemm_bf16_stats).  ``bn`` = (z, gamma, beta, mean, rstd) for stat_mode 2.  Returns (out, fused)."""
    import ctypes
    lib = _lib.load()
    assert x1.is_cuda and x1.dtype == torch.bfloat16 and x1.is_contiguous() and x1.dim() == 4
    B, H, W, c1 = x1.shape
    c2 = x2.shape[3] if x2 is not None else 0
    c_out, taps, c_in = weight_packed.shape
    assert weight_packed.dtype == torch.bfloat16 and weight_packed.is_contiguous() and c_in == c1 + c2
    assert sums.dtype == torch.float32 and sums.numel() >= 2 * c_out and sums.is_contiguous()
    out = torch.empty((B, H, W, c_out), dtype=torch.bfloat16, device=x1.device)
    z = gamma = beta = mean = rstd = None
    if stat_mode == 2:
        z, gamma, beta, mean, rstd = bn
        assert z.dtype == torch.bfloat16 and z.is_contiguous() and tuple(z.shape) == (B, H, W, c_out)
    fused = ctypes.c_int32(0)
    ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
    with torch.cuda.device(x1.device):
        rc = lib.im2im_conv_igemm_bf16_stats(x1.data_ptr(), c1, ptr(x2), c2, weight_packed.data_ptr(), B, H, W, c_out, taps,
                                             out.data_ptr(), stat_mode, sums.data_ptr(), ptr(z), ptr(gamma), ptr(beta),
                                             ptr(mean), ptr(rstd), ctypes.byref(fused),
                                             torch.cuda.current_stream(x1.device).cuda_stream)
    _lib.check(rc, "im2im_conv_igemm_bf16_stats")
    return out, bool(fused.value)


def conv_igemm_pool(x1: torch.Tensor, weight_packed: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
                    x2: Optional[torch.Tensor] = None):
    """conv_igemm (bf16 out) that also returns maxpool2x2 of its output when the layer runs on the halo kernel
    (im2im_conv_igemm_bf16_pool): (out, pooled or None).  ``pooled`` is bit-identical to maxpool2x2 of ``out``."""
    import ctypes
    lib = _lib.load()
    assert x1.is_cuda and x1.dtype == torch.bfloat16 and x1.is_contiguous() and x1.dim() == 4
    B, H, W, c1 = x1.shape
    c2 = x2.shape[3] if x2 is not None else 0
    c_out, taps, c_in = weight_packed.shape
    assert weight_packed.dtype == torch.bfloat16 and weight_packed.is_contiguous() and c_in == c1 + c2
    if H % 2 or W % 2:
        return conv_igemm(x1, weight_packed, bias, relu, x2), None
    out = torch.empty((B, H, W, c_out), dtype=torch.bfloat16, device=x1.device)
    pooled = torch.empty((B, H // 2, W // 2, c_out), dtype=torch.bfloat16, device=x1.device)
    done = ctypes.c_int32(0)
    with torch.cuda.device(x1.device):
        rc = lib.im2im_conv_igemm_bf16_pool(x1.data_ptr(), c1, x2.data_ptr() if x2 is not None else None, c2,
                                            weight_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
                                            B, H, W, c_out, taps, 1 if relu else 0, out.data_ptr(), pooled.data_ptr(),
                                            ctypes.byref(done), torch.cuda.current_stream(x1.device).cuda_stream)
    _lib.check(rc, "im2im_conv_igemm_bf16_pool")
    return out, (pooled if done.value else None)


def round_to_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 tensor rounded to nearest (ties away from zero, like cvt.rna.tf32.f32) onto the TF32 grid (10-bit mantissa)."""
    bits = t.detach().float().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_conv_weight_tf32(weight: torch.Tensor) -> torch.Tensor:
    """torch Conv2d weight [c_out, c_in, kh, kw] -> fp32 [c_out, kh*kw, c_in] (tap-major K), TF32-rounded, contiguous."""
    c_out, c_in, kh, kw = weight.shape
    return round_to_tf32(weight.detach().float().permute(0, 2, 3, 1).reshape(c_out, kh * kw, c_in).contiguous())


def conv_igemm_tf32(x1: torch.Tensor, weight_packed: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
                    x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Reference-precision convolution (im2im_conv_igemm_tf32): fp32 NHWC in, tcgen05 kind::tf32, fp32 NHWC out."""
    lib = _lib.load()
    assert x1.is_cuda and x1.dtype == torch.float32 and x1.is_contiguous() and x1.dim() == 4
    B, H, W, c1 = x1.shape
    c2 = 0
    if x2 is not None:
        assert x2.is_cuda and x2.dtype == torch.float32 and x2.is_contiguous() and tuple(x2.shape[:3]) == (B, H, W)
        c2 = x2.shape[3]
    c_out, taps, c_in = weight_packed.shape
    assert weight_packed.dtype == torch.float32 and weight_packed.is_contiguous() and c_in == c1 + c2
    out = torch.empty((B, H, W, c_out), dtype=torch.float32, device=x1.device)
    with torch.cuda.device(x1.device):
        rc = lib.im2im_conv_igemm_tf32(x1.data_ptr(), c1, x2.data_ptr() if x2 is not None else None, c2,
                                       weight_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
                                       B, H, W, c_out, taps, 1 if relu else 0, out.data_ptr(),
                                       torch.cuda.current_stream(x1.device).cuda_stream)
    _lib.check(rc, "im2im_conv_igemm_tf32")
    return out


def conv_wgrad(x: torch.Tensor, dz: torch.Tensor, taps: int = 9, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 dW [c_out, taps, c_in] (+= when ``out`` is given) from NHWC bf16 input ``x`` and output gradient ``dz``."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and dz.dtype == torch.bfloat16 and dz.is_contiguous()
    B, H, W, c_in = x.shape
    assert tuple(dz.shape[:3]) == (B, H, W)
    c_out = dz.shape[3]
    if out is None:
        out = torch.zeros((c_out, taps, c_in), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.im2im_conv_wgrad_bf16(x.data_ptr(), dz.data_ptr(), B, H, W, c_in, c_out, taps, out.data_ptr(),
                                       torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "im2im_conv_wgrad_bf16")
    return out


def pack_dgrad_weight(weight: torch.Tensor) -> torch.Tensor:
    """Weights for the data gradient: dX = conv3x3(dZ, W') with W'[ci, tap', co] = W[co, ci, flipped tap].

    torch weight [c_out, c_in, kh, kw] -> bf16 [c_in, kh*kw, c_out] with the taps reversed (180 degree rotation)."""
    c_out, c_in, kh, kw = weight.shape
    w = weight.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(c_in, kh * kw, c_out)
    return w.contiguous().to(torch.bfloat16)


def head_tc_applicable(H: int, W: int, n_out: int, c_mid: int) -> bool:
    """The tensor-core head (im2im_head_conv3x3_tc_f32) needs 8x16-pixel tiles, <= 32 output planes, <= 64 features."""
    return W % 8 == 0 and H % 16 == 0 and 1 <= n_out <= 32 and c_mid <= 64


def pad_head_weight(weight: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Stacked head weight fp32 [n_out, c_mid, 3, 3] -> fp32 [64, 64, 3, 3] with zero rows / input channels (``out`` is a
    persistent zero-initialised buffer: only the live block is rewritten)."""
    n_out, c_mid = weight.shape[:2]
    if out is None:
        out = torch.zeros((64, 64, 3, 3), dtype=torch.float32, device=weight.device)
    out[:n_out, :c_mid].copy_(weight.detach())
    return out


def head_conv_tc(x: torch.Tensor, weight_packed: torch.Tensor, bias: Optional[torch.Tensor], n_out: int,
                 act_kind: int = 0, act_from: int = 0, tap_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Head on tensor cores: x NHWC bf16 [B,H,W,64], weight bf16 [64,9,64] (pad_head_weight + pack_conv_weight) ->
    fp32 planes [B, n_out, H, W] with the head's activation fused.  ``tap_bias`` (fp32 [9, n_out]): the head has a 1x1
    convolution folded into it (fold_outconv_into_head): border pixels drop the bias of the taps in the padding."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[3] == 64
    assert weight_packed.dtype == torch.bfloat16 and weight_packed.is_contiguous() and tuple(weight_packed.shape) == (64, 9, 64)
    B, H, W, _ = x.shape
    out = torch.empty((B, n_out, H, W), dtype=torch.float32, device=x.device)
    st = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        if tap_bias is not None:
            assert tap_bias.is_cuda and tap_bias.dtype == torch.float32 and tap_bias.is_contiguous() and tuple(tap_bias.shape) == (9, n_out)
            rc = lib.im2im_head_conv3x3_tc_folded_f32(x.data_ptr(), weight_packed.data_ptr(), bias.data_ptr(),
                                                      tap_bias.data_ptr(), B, H, W, n_out, act_kind, act_from,
                                                      out.data_ptr(), st)
        else:
            rc = lib.im2im_head_conv3x3_tc_f32(x.data_ptr(), weight_packed.data_ptr(),
                                               bias.data_ptr() if bias is not None else None, B, H, W, n_out, act_kind,
                                               act_from, out.data_ptr(), st)
    _lib.check(rc, "im2im_head_conv3x3_tc_f32")
    return out


def fold_outconv_into_head(head_w: torch.Tensor, head_b: torch.Tensor, out_w: torch.Tensor, out_b: torch.Tensor):
    """head(OutConv(x)) as ONE 3x3 convolution of x: (weight fp32 [n_out, c_in, 3, 3], bias fp32 [n_out], tap_bias fp32
    [9 border classes, n_out]).  head_w [n_out, c_mid, 3, 3], out_w [c_mid, c_in, 1, 1] (unet_parts.py:87-93 followed by
    quantile_layer.py:15-20).  OutConv's bias reaches an output pixel once per in-range tap (its output is zero-PADDED, not
    bias-padded, at the image border): bias = head_b + sum over taps of tap_bias, and the kernel takes the out-of-range
    taps' share back on border pixels."""
    hw64, ow64 = head_w.double(), out_w.double()[:, :, 0, 0]
    w = torch.einsum('omyx,mc->ocyx', hw64, ow64).float().contiguous()
    tap = torch.einsum('omyx,m->oyx', hw64, out_b.double())                      # [n_out, 3, 3]
    bias = (head_b.double() + tap.sum((1, 2))).float().contiguous()
    # border class = 3 * (0 inside | 1 first row | 2 last row) + (0 inside | 1 first column | 2 last column)
    rows = (slice(0, 0), slice(0, 1), slice(2, 3))                               # kernel rows in the padding per row class
    border = torch.zeros(9, head_w.shape[0], dtype=torch.float64, device=head_w.device)
    for rc in range(3):
        for cc in range(3):
            oob = torch.zeros(3, 3, dtype=torch.bool, device=head_w.device)
            oob[rows[rc], :] = True
            oob[:, rows[cc]] = True
            border[3 * rc + cc] = (tap * oob).sum((1, 2))
    return w, bias, border.float().contiguous()


def head_conv_tc_hist(x: torch.Tensor, weight_packed: torch.Tensor, bias: Optional[torch.Tensor], labels: torch.Tensor,
                      lambdas_sorted: torch.Tensor, hist: torch.Tensor, act_kind: int = 0, act_from: int = 0,
                      out: Optional[torch.Tensor] = None, tap_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One-channel quantile head on tensor cores whose epilogue books every pixel's rank on the ascending lambda grid into
    ``hist`` (int32 [B, L+1], accumulated into) instead of writing the (B, 3, 1, H, W) head tensor
    (im2im_head_conv3x3_tc_hist).  ``labels`` fp32 [B, 1, H, W]; ``out`` (fp32 [B, 3, H, W]) also receives the planes."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[3] == 64
    assert weight_packed.dtype == torch.bfloat16 and weight_packed.is_contiguous() and tuple(weight_packed.shape) == (64, 9, 64)
    B, H, W, _ = x.shape
    L = lambdas_sorted.numel()
    assert labels.is_cuda and labels.dtype == torch.float32 and labels.is_contiguous() and labels.numel() == B * H * W
    assert lambdas_sorted.is_cuda and lambdas_sorted.dtype == torch.float32 and lambdas_sorted.is_contiguous()
    assert hist.is_cuda and hist.is_contiguous() and hist.element_size() == 4 and tuple(hist.shape) == (B, L + 1)
    if out is not None:
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, 3, H, W)
    with torch.cuda.device(x.device):
        rc = lib.im2im_head_conv3x3_tc_hist(x.data_ptr(), weight_packed.data_ptr(),
                                            bias.data_ptr() if bias is not None else None,
                                            tap_bias.data_ptr() if tap_bias is not None else None, B, H, W, 3, act_kind, act_from,
                                            out.data_ptr() if out is not None else None, labels.data_ptr(),
                                            lambdas_sorted.data_ptr(), L, hist.data_ptr(),
                                            torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "im2im_head_conv3x3_tc_hist")
    return hist


def planar_to_nhwc64(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 planes [B, n, H, W] -> bf16 NHWC [B, H, W, 64], channels >= n zero.

    ``out``: a persistent [B, H, W, 64] bf16 buffer whose channels 8..63 are zero and stay zero (n <= 8): only channels
    0..7 are rewritten - an eighth of the bytes."""
    lib = _lib.load()
    assert src.is_cuda and src.dtype == torch.float32 and src.is_contiguous() and src.dim() == 4 and src.shape[1] <= 64
    B, n, H, W = src.shape
    with torch.cuda.device(src.device):
        if out is not None and n <= 8:
            assert out.dtype == torch.bfloat16 and out.is_contiguous() and tuple(out.shape) == (B, H, W, 64)
            rc = lib.im2im_planar_to_nhwc64_first8_bf16(src.data_ptr(), n, B, H, W, out.data_ptr(),
                                                        torch.cuda.current_stream(src.device).cuda_stream)
            _lib.check(rc, "im2im_planar_to_nhwc64_first8_bf16")
            return out
        dst = torch.empty((B, H, W, 64), dtype=torch.bfloat16, device=src.device)
        rc = lib.im2im_planar_to_nhwc64_bf16(src.data_ptr(), n, B, H, W, dst.data_ptr(),
                                             torch.cuda.current_stream(src.device).cuda_stream)
    _lib.check(rc, "im2im_planar_to_nhwc64_bf16")
    return dst
