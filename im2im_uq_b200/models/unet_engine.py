"""Native sm_100a inference forward of UNet + quantile head (Path B, inference half).

Replaces, for ``model.eval()`` forwards on CUDA under ``torch.no_grad()``, the library calls behind
``ModelWithUncertainty.forward`` (core/models/add_uncertainty.py:25-27) -> ``UNet.forward``
(core/models/trunks/unet.py:33-46) -> ``QuantileRegressionLayer.forward`` (core/models/finallayers/quantile_layer.py:19-21):
BatchNorm (running statistics) is folded into each convolution's weight/bias, activations are NHWC bf16, the 3x3 / 1x1
convolutions run as implicit GEMMs on tcgen05 (fp32 accumulation), ReLU is fused, the skip concatenation is never
materialised, and the head writes the reference's (B, 3, C_out, H, W) fp32 tensor directly.

Two precisions (``model.native_precision`` or the environment variable IM2IM_UNET_PRECISION):
  "bf16" (default)  bf16 operands / activations, tcgen05 kind::f16 - the fast mode (outputs within ~1e-2 of fp32)
  "tf32"            the reference's precision: its modules are fp32 (unet_parts.py:16-21) and torch runs their cuDNN
                    convolutions as TF32 on a GPU; fp32 NHWC activations, tcgen05 kind::tf32 (rel-L2 ~5e-4 vs fp32),
                    about half the throughput.  SURVEY.md §7 hard part 3.
"""
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from ..conv import (conv_igemm, conv_igemm_pool, conv_igemm_tf32, fold_outconv_into_head, head_conv_tc, head_conv_tc_hist,
                    head_tc_applicable, pack_conv_weight,
                    pack_conv_weight_tf32, pad_head_weight)


def _fold_bn(conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d]) -> Tuple[torch.Tensor, torch.Tensor]:
    """conv -> BN(eval) == conv with w*s and (b-mean)*s+beta, s = gamma/sqrt(var+eps) (fp32)."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is not None:
        s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * s[:, None, None, None]
        b = (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return w.contiguous(), b.contiguous()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class UNetInferenceEngine:
    """Folded/packed weights + the launch sequence of the native forward for one ModelWithUncertainty."""

    def __init__(self, model):
        self.model = model
        self._stamp = None
        self.refresh()

    def _precision(self) -> str:
        p = getattr(self.model, "native_precision", None) or os.environ.get("IM2IM_UNET_PRECISION", "bf16")
        if p not in ("bf16", "tf32"):
            raise ValueError(f"native_precision must be 'bf16' or 'tf32', got {p!r}")
        return p

    # ---- weights
    def _param_stamp(self):
        # _lib.weights_generation(): raw-pointer / graph-replay writes (FusedAdam, bn_finalize, GraphedTrainStep) that
        # tensor._version cannot see
        return (_lib.weights_generation(), self._precision()) + tuple((p.data_ptr(), p._version) for p in self.model.parameters()) + \
            tuple((b.data_ptr(), b._version) for n, b in self.model.named_buffers() if b is not None and n != "lhat")

    def refresh(self):
        trunk, head = self.model.baseModel, self.model.last_layer
        self.precision = self._precision()
        tf32 = self.precision == "tf32"

        def double(dc):
            seq = dc.double_conv
            return _fold_bn(seq[0], seq[1]), _fold_bn(seq[3], seq[4])

        def packed(wb):
            w, b = wb
            return (pack_conv_weight_tf32(w) if tf32 else pack_conv_weight(w)), b

        (w0, b0), second = double(trunk.inc)
        self.first = (w0.contiguous(), b0)  # fp32 [64, c_in, 3, 3]
        self.inc2 = packed(second)
        self.down = []
        for blk in (trunk.down1, trunk.down2, trunk.down3, trunk.down4):
            a, b = double(blk.maxpool_conv[1])
            self.down.append((packed(a), packed(b)))
        self.up = []
        for blk in (trunk.up1, trunk.up2, trunk.up3, trunk.up4):
            a, b = double(blk.conv)
            self.up.append((packed(a), packed(b)))
        self.out = packed(_fold_bn(trunk.out.conv, None))
        planes = head_plane_convs(head)
        if planes is None or len(planes[0]) * planes[0][0].weight.shape[0] not in (2, 3, 4, 6, 9):
            self.head = None            # e.g. the 50-plane softmax head: native trunk, library head conv (forward())
            self._stamp = self._param_stamp()
            return
        convs, act, act_from = planes
        hw = torch.cat([c.weight for c in convs], dim=0).detach().float().contiguous()
        hb = torch.cat([c.bias for c in convs], dim=0).detach().float().contiguous()
        self.head = (hw, hb)
        self.n_planes = len(convs)
        self.c_out = convs[0].weight.shape[0]
        self.head_act = ({None: 0, "relu": 1, "abs": 2}[act], act_from * self.c_out)
        # tensor-core head (8x16-pixel tiles): OutConv padded to 64 output channels (rows >= 32 zero) feeds a 64 -> 64 halo
        # convolution whose first n_out outputs are the head's planes
        ow, ob = _fold_bn(trunk.out.conv, None)
        c_mid = ow.shape[0]
        self.c_mid = c_mid
        if not tf32 and c_mid <= 64 and hw.shape[0] <= 32 and ow.shape[1] % 64 == 0:
            ow64 = torch.zeros((64,) + tuple(ow.shape[1:]), dtype=torch.float32, device=ow.device)
            ob64 = torch.zeros(64, dtype=torch.float32, device=ow.device)
            ow64[:c_mid] = ow
            ob64[:c_mid] = ob
            self.out64 = (pack_conv_weight(ow64), ob64)
            self.head_tc = (pack_conv_weight(pad_head_weight(hw)), hb)
            # OutConv (1x1) folded into the head's 3x3 weights: one convolution of the 64-channel feature map, the
            # (B, H, W, 32) tensor in between is never written (IM2IM_NO_HEAD_FOLD=1 keeps the two convolutions)
            self.head_fold = None
            if (ow.shape[1] == 64 and tuple(ow.shape[2:]) == (1, 1) and hw.shape[0] <= 7
                    and not os.environ.get("IM2IM_NO_HEAD_FOLD")):
                wf, bf, tb = fold_outconv_into_head(hw, hb, ow, ob)
                self.head_fold = (pack_conv_weight(pad_head_weight(wf)), bf, tb)
        else:
            self.out64 = self.head_tc = self.head_fold = None
        self._stamp = self._param_stamp()

    # ---- single launches
    @staticmethod
    def _conv_first(x, w, b):
        lib = _lib.load()
        B, c_in, H, W = x.shape
        c_out = w.shape[0]
        y = torch.empty((B, H, W, c_out), dtype=torch.bfloat16, device=x.device)
        _lib.check(lib.im2im_conv_first_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), B, c_in, H, W, c_out, 1,
                                             y.data_ptr(), _stream(x.device)), "im2im_conv_first_bf16")
        return y

    @staticmethod
    def _pool(x):
        lib = _lib.load()
        B, H, W, C = x.shape
        y = torch.empty((B, H // 2, W // 2, C), dtype=torch.bfloat16, device=x.device)
        _lib.check(lib.im2im_maxpool2x2_bf16(x.data_ptr(), B, H, W, C, y.data_ptr(), _stream(x.device)),
                   "im2im_maxpool2x2_bf16")
        return y

    @staticmethod
    def _upsample_to(x, H_out, W_out):
        lib = _lib.load()
        B, h, w, C = x.shape
        y = torch.empty((B, H_out, W_out, C), dtype=torch.bfloat16, device=x.device)
        _lib.check(lib.im2im_upsample2x_bilinear_bf16(x.data_ptr(), B, h, w, C, H_out, W_out, y.data_ptr(),
                                                      _stream(x.device)), "im2im_upsample2x_bilinear_bf16")
        return y

    def _head(self, x):
        lib = _lib.load()
        B, H, W, C = x.shape
        hw, hb = self.head
        n_out = hw.shape[0]
        y = torch.empty((B, n_out, H, W), dtype=torch.float32, device=x.device)
        _lib.check(lib.im2im_head_conv3x3_act_f32(x.data_ptr(), hw.data_ptr(), hb.data_ptr(), None, B, H, W, C, C, n_out,
                                                  self.head_act[0], self.head_act[1], y.data_ptr(), _stream(x.device)),
                   "im2im_head_conv3x3_act_f32")
        return y.view(B, self.n_planes, self.c_out, H, W)

    def _trunk(self, x: torch.Tensor) -> torch.Tensor:
        """UNet body up to the last 64-channel feature map (NHWC bf16), the input of OutConv (unet.py:33-45)."""
        a = self._conv_first(x, *self.first)
        # a skip layer's convolution also writes its own 2x2 max-pool where it runs on the halo kernel (the pooling pass over
        # the full-resolution skip tensor disappears); deeper layers pool separately
        s, pooled = conv_igemm_pool(a, self.inc2[0], self.inc2[1], relu=True)
        skips: List[torch.Tensor] = [s]
        for k, (c1, c2) in enumerate(self.down):
            p = pooled if pooled is not None else self._pool(skips[-1])
            p = conv_igemm(p, c1[0], c1[1], relu=True)
            if k + 1 < len(self.down):
                s, pooled = conv_igemm_pool(p, c2[0], c2[1], relu=True)
            else:
                s, pooled = conv_igemm(p, c2[0], c2[1], relu=True), None
            skips.append(s)
        y = skips.pop()
        for (c1, c2) in self.up:
            skip = skips.pop()
            u = self._upsample_to(y, skip.shape[1], skip.shape[2])
            y = conv_igemm(skip, c1[0], c1[1], relu=True, x2=u)   # torch.cat([skip, up]) without the copy
            y = conv_igemm(y, c2[0], c2[1], relu=True)
        return y

    def hist_applicable(self, x: torch.Tensor) -> bool:
        """Can ``forward_hist`` take this batch?  (bf16 mode, one-channel QuantileRegressionLayer, 8x16-pixel tiles.)"""
        from .quantile_layer import QuantileRegressionLayer
        if self._stamp != self._param_stamp():
            self.refresh()
        return (self.precision == "bf16" and type(self.model.last_layer) is QuantileRegressionLayer
                and self.head is not None and self.head_tc is not None and self.n_planes == 3 and self.c_out == 1
                and self.head_act[0] == 0 and x.is_cuda and x.dim() == 4
                and head_tc_applicable(x.shape[2], x.shape[3], 3, self.c_mid))

    def forward_hist(self, x: torch.Tensor, labels: torch.Tensor, lambdas_sorted: torch.Tensor, hist: torch.Tensor,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The forward of ``forward`` with the head's epilogue booking every pixel's RCPS rank into ``hist`` (int32
        [B, L+1]) instead of writing the (B, 3, 1, H, W) tensor: calibrate_model.py:121-123 + :134-136 of the reference
        without the 12 bytes per pixel in between.  ``out`` (fp32 [B, 3, H, W]) additionally receives the planes."""
        if not self.hist_applicable(x):
            raise _lib.Im2ImError("forward_hist: needs the bf16 engine, a one-channel quantile head and H % 16 == W % 8 == 0")
        x = x.contiguous().float()
        with torch.cuda.device(x.device):
            y = self._trunk(x)
            if self.head_fold is not None:
                return head_conv_tc_hist(y, self.head_fold[0], self.head_fold[1], labels, lambdas_sorted, hist, out=out,
                                         tap_bias=self.head_fold[2])
            m = conv_igemm(y, self.out64[0], self.out64[1], relu=False)
            return head_conv_tc_hist(m, self.head_tc[0], self.head_tc[1], labels, lambdas_sorted, hist, out=out)

    # ---- the forward
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._stamp != self._param_stamp():
            self.refresh()
        if not x.is_cuda:
            raise _lib.Im2ImError("native UNet forward needs a CUDA tensor")
        x = x.contiguous().float()
        if self.precision == "tf32":
            return self._forward_tf32(x)
        with torch.cuda.device(x.device):
            y = self._trunk(x)
            if (self.head is not None and self.head_tc is not None
                    and head_tc_applicable(y.shape[1], y.shape[2], self.head[0].shape[0], self.c_mid)):
                n_out = self.head[0].shape[0]
                if self.head_fold is not None:   # OutConv folded into the head: one 3x3 convolution of y
                    out = head_conv_tc(y, self.head_fold[0], self.head_fold[1], n_out, self.head_act[0], self.head_act[1],
                                       tap_bias=self.head_fold[2])
                else:
                    m = conv_igemm(y, self.out64[0], self.out64[1], relu=False)   # 1x1 OutConv, 64 -> 32 (+32 zero channels)
                    out = head_conv_tc(m, self.head_tc[0], self.head_tc[1], n_out, self.head_act[0], self.head_act[1])
                return out.view(x.shape[0], self.n_planes, self.c_out, y.shape[1], y.shape[2])
            m = conv_igemm(y, self.out[0], self.out[1], relu=False)  # 1x1 OutConv, 64 -> 32 (tensor cores)
            if self.head is None:  # heads without a native kernel (softmax: 50 planes) run their own module on the features
                return self.model.last_layer(m.permute(0, 3, 1, 2).float())
            return self._head(m)


def _forward_tf32(self, x: torch.Tensor) -> torch.Tensor:
    """The same launch sequence on fp32 NHWC activations with kind::tf32 convolutions (the reference's precision)."""
    lib = _lib.load()
    dev = x.device
    st = _stream(dev)
    B, c_in, H, W = x.shape

    def first(x):
        w, b = self.first
        y = torch.empty((B, H, W, w.shape[0]), dtype=torch.float32, device=dev)
        _lib.check(lib.im2im_conv_first_nhwc_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), B, c_in, H, W, w.shape[0], 1,
                                                 y.data_ptr(), st), "im2im_conv_first_nhwc_f32")
        return y

    def pool(t):
        b_, h, w_, c = t.shape
        y = torch.empty((b_, h // 2, w_ // 2, c), dtype=torch.float32, device=dev)
        _lib.check(lib.im2im_maxpool2x2_nhwc_f32(t.data_ptr(), b_, h, w_, c, y.data_ptr(), st), "im2im_maxpool2x2_nhwc_f32")
        return y

    def upsample(t, ho, wo):
        b_, h, w_, c = t.shape
        y = torch.empty((b_, ho, wo, c), dtype=torch.float32, device=dev)
        _lib.check(lib.im2im_upsample2x_bilinear_nhwc_f32(t.data_ptr(), b_, h, w_, c, ho, wo, y.data_ptr(), st),
                   "im2im_upsample2x_bilinear_nhwc_f32")
        return y

    with torch.cuda.device(dev):
        skips: List[torch.Tensor] = [conv_igemm_tf32(first(x), self.inc2[0], self.inc2[1], relu=True)]
        for (c1, c2) in self.down:
            p = conv_igemm_tf32(pool(skips[-1]), c1[0], c1[1], relu=True)
            skips.append(conv_igemm_tf32(p, c2[0], c2[1], relu=True))
        y = skips.pop()
        for (c1, c2) in self.up:
            skip = skips.pop()
            u = upsample(y, skip.shape[1], skip.shape[2])
            y = conv_igemm_tf32(skip, c1[0], c1[1], relu=True, x2=u)
            y = conv_igemm_tf32(y, c2[0], c2[1], relu=True)
        m = conv_igemm_tf32(y, self.out[0], self.out[1], relu=False)           # 1x1 OutConv, 64 -> 32
        if self.head is None:
            return self.model.last_layer(m.permute(0, 3, 1, 2))
        hw, hb = self.head
        n_out, c_mid = hw.shape[0], m.shape[3]
        out = torch.empty((B, n_out, H, W), dtype=torch.float32, device=dev)
        _lib.check(lib.im2im_head_conv3x3_act_nhwc_f32(m.data_ptr(), hw.data_ptr(), hb.data_ptr(), None, B, H, W, c_mid, c_mid,
                                                       n_out, self.head_act[0], self.head_act[1], out.data_ptr(), st),
                   "im2im_head_conv3x3_act_nhwc_f32")
        return out.view(B, self.n_planes, self.c_out, H, W)


UNetInferenceEngine._forward_tf32 = _forward_tf32


def head_plane_convs(head):
    """(convs in plane order, activation name or None, first activated plane) of a stacked-planes head, else None.

    Covers QuantileRegressionLayer (quantile_layer.py:15-20) and the heads of heads.py that stack one 3x3 conv per
    plane (gaussian: relu on the variance plane; residual magnitude: abs on the magnitude plane)."""
    from .heads import _StackedPlanesHead
    from .quantile_layer import QuantileRegressionLayer
    if type(head) is QuantileRegressionLayer:
        return [head.lower, head.prediction, head.upper], None, 0
    if isinstance(head, _StackedPlanesHead):
        return head.plane_convs(), head.act, head.act_from
    return None


def native_forward_applicable(model, x) -> bool:
    from .unet import UNet
    if not (not model.training and torch.is_tensor(x) and x.is_cuda and not torch.is_grad_enabled()
            and type(model.baseModel) is UNet and model.baseModel.bilinear
            and getattr(model, "use_native_inference", True)
            and x.dim() == 4 and x.shape[1] <= 8 and min(x.shape[2], x.shape[3]) >= 16):
        return False
    if not isinstance(model.last_layer, nn.Module):
        return False
    return True
