"""Native sm_100a training step of UNet + quantile head (Path B, training half).

Replaces the library calls behind the reference's hot training loop (core/scripts/train.py:147-165):
    labels_pred = net(*x)                    -> UNetTrainEngine.forward   (tcgen05 convs, BatchNorm batch statistics)
    loss = net.loss_fn(labels_pred, labels)  -> fused pinball+MSE kernel  (quantile_layer.py:23-32)
    loss.backward()                          -> UNetTrainEngine.backward  (dgrad = igemm with flipped weights, wgrad =
                                                tcgen05 MN-major GEMM over pixels, BN/ReLU/pool/upsample backward)
    optimizer.step()                         -> FusedAdam (one kernel over the flat fp32 parameter buffer)
Data-parallel: one process per GPU, per-replica BatchNorm statistics (nn.DataParallel semantics, train.py:112-115),
ONE NCCL all-reduce of the flat fp32 gradient buffer (69 MB) per step.

The engine plugs in behind ``ModelWithUncertainty.forward`` through a torch.autograd.Function, so the reference's own
loop (forward, loss_fn, backward, optimizer.step) runs unchanged.

Precision: activations and GEMM operands bf16, accumulation / statistics / parameters / gradients fp32.
Known deviation: convolution biases that feed a BatchNorm receive an exactly-zero gradient (their true gradient is
zero up to rounding noise, because the batch mean cancels them); torch produces ~1e-9 noise there instead.
"""
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import _lib
from ..conv import (conv_igemm, conv_igemm_stats, conv_wgrad, head_conv_tc, head_tc_applicable, pack_conv_weight,
                    pack_dgrad_weight, planar_to_nhwc64)


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return t.data_ptr() if t is not None else None


class _ConvBN:
    """conv3x3 (tensor cores) -> BatchNorm(batch stats) -> ReLU, with what backward needs."""

    def __init__(self, conv: nn.Conv2d, bn: nn.BatchNorm2d):
        self.conv, self.bn = conv, bn
        self.c_out, self.c_in = conv.weight.shape[0], conv.weight.shape[1]

    def pack(self):
        """fp32 master weight -> the two bf16 GEMM operands, one kernel (buffers are reused across steps)."""
        w = self.conv.weight.detach()
        if getattr(self, "w_fwd", None) is None or self.w_fwd.device != w.device:
            self.w_fwd = torch.empty((self.c_out, 9, self.c_in), dtype=torch.bfloat16, device=w.device)
            self.w_bwd = torch.empty((self.c_in, 9, self.c_out), dtype=torch.bfloat16, device=w.device)
        wc = w if w.is_contiguous() else w.contiguous()
        with torch.cuda.device(w.device):
            _lib.check(_lib.load().im2im_pack_conv_weights(wc.data_ptr(), self.c_out, self.c_in, 9, self.w_fwd.data_ptr(),
                                                           self.w_bwd.data_ptr(), _st(w.device)), "pack_conv_weights")


class _ZeroPool:
    """One zero-filled fp32 buffer per pass handed out in slices: a single fill kernel instead of one per small
    accumulator (BatchNorm sums, weight-gradient buffers, ...) - the step is made of ~400 launches, the tiny ones count."""

    def __init__(self, n_floats: int, device):
        self.buf = torch.zeros(max(n_floats, 4), dtype=torch.float32, device=device)
        self.off = 0

    def take(self, *shape):
        k = 1
        for d in shape:
            k *= d
        assert self.off + k <= self.buf.numel(), "zero pool exhausted"
        v = self.buf[self.off:self.off + k].view(*shape)
        self.off += (k + 3) // 4 * 4   # keep every slice 16-byte aligned
        return v

    @staticmethod
    def padded(k: int) -> int:
        return (k + 3) // 4 * 4


class UNetTrainEngine:
    def __init__(self, model):
        from .unet import UNet
        from .unet_engine import head_plane_convs
        assert type(model.baseModel) is UNet and head_plane_convs(model.last_layer) is not None
        self.model = model
        t, self.head = model.baseModel, model.last_layer
        self.lib = _lib.load()
        import os
        self.fuse_pool = os.environ.get("IM2IM_NO_POOL_FUSION") is None    # A/B switch (tools/train_profile.py)

        def dc(d):
            s = d.double_conv
            return _ConvBN(s[0], s[1]), _ConvBN(s[3], s[4])

        self.inc = dc(t.inc)
        self.down = [dc(b.maxpool_conv[1]) for b in (t.down1, t.down2, t.down3, t.down4)]
        self.up = [dc(b.conv) for b in (t.up1, t.up2, t.up3, t.up4)]
        self.out_conv = t.out.conv
        self.c_mid = self.out_conv.weight.shape[0]      # 32
        self.head_convs, self.head_act, act_from = head_plane_convs(self.head)   # plane order, relu/abs/None
        self.c_head = self.head_convs[0].weight.shape[0]
        self.n_planes = len(self.head_convs)
        self.n_out = self.n_planes * self.c_head
        self.act_from = act_from * self.c_head
        layers = [l for pair in [self.inc] + self.down + self.up for l in pair]
        pad = _ZeroPool.padded
        self._fwd_zero_floats = sum(pad(2 * l.c_out) for l in layers)
        self._bwd_zero_floats = (sum(2 * pad(l.c_out * 9 * l.c_in) for l in layers[1:])   # concat layers: two buffers
                                 + sum(pad(2 * l.c_out) for l in layers)   # BatchNorm-backward sums fused into dgrad epilogues
                                 + pad(64 * 64) + pad(self.n_out * self.c_mid * 9) + pad(self.n_out) + pad(128)
                                 + pad(64 * 9 * 64)
                                 + pad(layers[0].c_out * 8 * 9))

    # ------------------------------------------------------------------------------------------- primitive launches
    def _conv_bn_relu(self, x: torch.Tensor, layer: _ConvBN, saved: dict, x2: Optional[torch.Tensor] = None,
                      pool: bool = False):
        """conv3x3 on tensor cores -> BatchNorm (batch statistics) -> ReLU.  Where the layer runs on the halo kernel the
        statistics come out of the convolution's epilogue (im2im_conv_igemm_bf16_stats): no separate pass over z."""
        zero_pool = self.__dict__.get("_zero")
        C = layer.c_out
        sums = zero_pool.take(2 * C) if zero_pool is not None else torch.zeros(2 * C, dtype=torch.float32, device=x.device)
        if getattr(self, "fuse_stats", True):
            z, fused = conv_igemm_stats(x, layer.w_fwd, 1, sums, x2=x2)
        else:
            z, fused = conv_igemm(x, layer.w_fwd, x2=x2), False
        return self._bn_relu(z, layer, saved, sums=sums, have_stats=fused, pool=pool)

    def _bn_relu(self, z: torch.Tensor, layer: _ConvBN, saved: dict, sums: Optional[torch.Tensor] = None,
                 have_stats: bool = False, pool: bool = False):
        """BatchNorm (batch statistics) + ReLU of the bias-free convolution output z.  ``pool=True`` (skip layers): also
        returns maxpool2x2(y), computed in the same pass when H and W are even."""
        lib, dev = self.lib, z.device
        B, H, W, C = z.shape
        n_pix = B * H * W
        bn = layer.bn
        zero_pool = self.__dict__.get("_zero")     # set by forward(); a direct call (unit tests) allocates its own
        if sums is None:
            sums = zero_pool.take(2 * C) if zero_pool is not None else torch.zeros(2 * C, dtype=torch.float32, device=dev)
        scale = torch.empty(C, dtype=torch.float32, device=dev)
        shift = torch.empty_like(scale)
        rstd = torch.empty_like(scale)
        mean = torch.empty_like(scale)
        if not have_stats:
            _lib.check(lib.im2im_channel_stats_bf16(z.data_ptr(), n_pix, C, sums.data_ptr(), _st(dev)), "channel_stats")
        momentum = bn.momentum if bn.momentum is not None else 0.1
        track = bn.track_running_stats and bn.running_mean is not None
        _lib.check(lib.im2im_bn_finalize(sums.data_ptr(), n_pix, _ptr(layer.conv.bias), bn.weight.data_ptr(),
                                         bn.bias.data_ptr(), bn.eps, momentum, C,
                                         _ptr(bn.running_mean) if track else None,
                                         _ptr(bn.running_var) if track else None, scale.data_ptr(), shift.data_ptr(),
                                         mean.data_ptr(), rstd.data_ptr(), _st(dev)), "bn_finalize")
        if track and bn.num_batches_tracked is not None:
            if self.__dict__.get("_tracked") is not None:
                self._tracked.append(bn.num_batches_tracked)   # advanced together at the end of forward (one launch)
            else:
                bn.num_batches_tracked.add_(1)
        y = torch.empty_like(z)
        saved["z"], saved["mean"], saved["rstd"] = z, mean, rstd  # backward needs xhat everywhere: keep z, not y
        if pool and H % 2 == 0 and W % 2 == 0 and getattr(self, "fuse_pool", True):
            p = torch.empty((B, H // 2, W // 2, C), dtype=torch.bfloat16, device=dev)
            _lib.check(lib.im2im_bn_apply_relu_pool_bf16(z.data_ptr(), scale.data_ptr(), shift.data_ptr(), B, H, W, C,
                                                         y.data_ptr(), p.data_ptr(), _st(dev)), "bn_apply_relu_pool")
            return y, p
        _lib.check(lib.im2im_bn_apply_relu_bf16(z.data_ptr(), scale.data_ptr(), shift.data_ptr(), n_pix, C,
                                                y.data_ptr(), _st(dev)), "bn_apply_relu")
        return (y, self._pool(y)) if pool else y

    def _bn_relu_bwd(self, dy: torch.Tensor, layer: _ConvBN, saved: dict, grads: Dict):
        lib, dev = self.lib, dy.device
        z = saved["z"]
        B, H, W, C = z.shape
        sums = torch.empty(2 * C, dtype=torch.float32, device=dev)
        dz = torch.empty_like(z)
        _lib.check(lib.im2im_bn_relu_bwd_bf16(dy.data_ptr(), z.data_ptr(), layer.bn.weight.data_ptr(),
                                              layer.bn.bias.data_ptr(), saved["mean"].data_ptr(),
                                              saved["rstd"].data_ptr(), B * H * W, C, sums.data_ptr(), dz.data_ptr(),
                                              _st(dev)), "bn_relu_bwd")
        grads[layer.bn.bias] = sums[:C]
        grads[layer.bn.weight] = sums[C:]
        # layer.conv.bias gets no entry: its gradient is exactly zero (cancelled by the batch mean, see module doc), and
        # returning None to autograd leaves the zeroed .grad untouched instead of launching a fill and an add per layer
        return dz

    def _bn_relu_pool_bwd(self, d_skip: torch.Tensor, d_p: torch.Tensor, layer: _ConvBN, saved: dict, grads: Dict):
        """BatchNorm+ReLU backward of a skip layer whose output also fed a 2x2 max-pool: d_skip (gradient through the skip
        connection) and d_p (gradient of the pooled tensor) in, dz out - the pool's scatter pass is folded in."""
        z = saved["z"]
        B, H, W, C = z.shape
        dev = z.device
        if not (H % 2 == 0 and W % 2 == 0 and getattr(self, "fuse_pool", True)):
            _lib.check(self.lib.im2im_maxpool2x2_bwd_bf16(saved["y_out"].data_ptr(), d_p.data_ptr(), B, H, W, C, 1,
                                                          d_skip.data_ptr(), _st(dev)), "maxpool_bwd")
            return self._bn_relu_bwd(d_skip, layer, saved, grads)
        sums = torch.empty(2 * C, dtype=torch.float32, device=dev)
        dz = torch.empty_like(z)
        _lib.check(self.lib.im2im_bn_relu_pool_bwd_bf16(d_skip.data_ptr(), d_p.data_ptr(), z.data_ptr(),
                                                        layer.bn.weight.data_ptr(), layer.bn.bias.data_ptr(),
                                                        saved["mean"].data_ptr(), saved["rstd"].data_ptr(), B, H, W, C,
                                                        sums.data_ptr(), dz.data_ptr(), _st(dev)), "bn_relu_pool_bwd")
        grads[layer.bn.bias] = sums[:C]
        grads[layer.bn.weight] = sums[C:]
        return dz

    def _side_stream(self, dev):
        st = self.__dict__.get("_side")
        if st is None or st.device != dev:
            st = torch.cuda.Stream(device=dev)
            self._side = st
        return st

    @staticmethod
    def _w_to_torch(dw: torch.Tensor, c_in: int):
        c_out, taps, _ = dw.shape
        k = 3 if taps == 9 else 1
        return dw.view(c_out, k, k, c_in).permute(0, 3, 1, 2)

    def _pool(self, x):
        B, H, W, C = x.shape
        y = torch.empty((B, H // 2, W // 2, C), dtype=torch.bfloat16, device=x.device)
        _lib.check(self.lib.im2im_maxpool2x2_bf16(x.data_ptr(), B, H, W, C, y.data_ptr(), _st(x.device)), "maxpool")
        return y

    def _upsample_to(self, x, Ho, Wo):
        B, h, w, C = x.shape
        y = torch.empty((B, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
        _lib.check(self.lib.im2im_upsample2x_bilinear_bf16(x.data_ptr(), B, h, w, C, Ho, Wo, y.data_ptr(),
                                                           _st(x.device)), "upsample")
        return y

    # ------------------------------------------------------------------------------------------- forward
    def forward(self, x: torch.Tensor):
        lib, dev = self.lib, x.device
        x = x.contiguous().float()
        B, c_in, H, W = x.shape
        for a, b in [self.inc] + self.down + self.up:
            b.pack()
            if a is not self.inc[0]:
                a.pack()
        ctx = {"x": x, "layers": []}
        _lib.weights_changed()          # bn_finalize updates the running statistics through raw pointers
        self._zero = _ZeroPool(self._fwd_zero_floats, dev)
        self._tracked = []
        with torch.cuda.device(dev):
            # first conv (CUDA cores), no bias (cancelled by BN), no ReLU: z0
            first = self.inc[0]
            w0 = first.conv.weight.detach().float().contiguous()
            z = torch.empty((B, H, W, first.c_out), dtype=torch.bfloat16, device=dev)
            _lib.check(lib.im2im_conv_first_bf16(x.data_ptr(), w0.data_ptr(), None, B, c_in, H, W, first.c_out, 0,
                                                 z.data_ptr(), _st(dev)), "conv_first")
            s0 = {}
            y0 = self._bn_relu(z, first, s0)
            s1 = {"x_in": y0}
            x1, p = self._conv_bn_relu(y0, self.inc[1], s1, pool=True)    # x1 feeds the skip connection and down1's pool
            s1["y_out"] = x1
            ctx["inc"] = (s0, s1)
            skips = [x1]
            ctx["down"] = []
            for k, (a, b) in enumerate(self.down):
                sa = {"x_in": p}
                ya = self._conv_bn_relu(p, a, sa)
                sb = {"x_in": ya}
                if k < len(self.down) - 1:
                    xk, p = self._conv_bn_relu(ya, b, sb, pool=True)
                    sb["y_out"] = xk
                else:
                    xk = self._conv_bn_relu(ya, b, sb)                  # x5: bottom of the U, not pooled
                skips.append(xk)
                ctx["down"].append((sa, sb))
            ctx["skips"] = list(skips)
            y = skips.pop()
            ctx["up"] = []
            for (a, b) in self.up:
                skip = skips.pop()
                u = self._upsample_to(y, skip.shape[1], skip.shape[2])
                sa = {"x_in": skip, "x_in2": u, "low_shape": tuple(y.shape)}
                ya = self._conv_bn_relu(skip, a, sa, x2=u)
                sb = {"x_in": ya}
                y = self._conv_bn_relu(ya, b, sb)
                ctx["up"].append((sa, sb))
            # 1x1 out conv, output channels zero-padded 32 -> 64 so the result feeds the 64-channel kernels
            w_out = self.out_conv.weight.detach()
            c_feat = w_out.shape[1]
            pads = self.__dict__.get("_out_pad")
            if pads is None or pads[0].device != dev:   # rows c_mid..63 stay zero for the engine's lifetime
                pads = (torch.zeros((64, 1, c_feat), dtype=torch.bfloat16, device=dev),
                        torch.zeros(64, dtype=torch.float32, device=dev))
                self._out_pad = pads
            w_pad, b_pad = pads
            w_pad[:self.c_mid, 0] = w_out.view(self.c_mid, c_feat)          # fp32 -> bf16 in the copy
            b_pad[:self.c_mid] = self.out_conv.bias.detach()
            m = conv_igemm(y, w_pad, b_pad, relu=False)
            ctx["y_last"], ctx["m"], ctx["w_out_pad"] = y, m, w_pad
            # head (CUDA cores)
            hw_ = torch.cat([c.weight for c in self.head_convs], 0).detach().float().contiguous()
            hb = torch.cat([c.bias for c in self.head_convs], 0).detach().float().contiguous()
            ctx["head_tc"] = head_tc_applicable(H, W, self.n_out, self.c_mid)
            if ctx["head_tc"]:
                # head on tensor cores: the stacked 3x3 convs as one 64 -> 64 halo convolution over the zero-padded features
                # (weights bf16 [64, 9, 64], rows >= n_out and input channels >= c_mid zero), planes written directly
                bufs = self.__dict__.get("_head_bufs")
                if bufs is None or bufs[0].device != dev:
                    bufs = (torch.zeros((64, 64, 3, 3), dtype=torch.float32, device=dev),
                            torch.empty((64, 9, 64), dtype=torch.bfloat16, device=dev),
                            torch.empty((64, 9, 64), dtype=torch.bfloat16, device=dev))
                    self._head_bufs = bufs
                w64, w_fwd, w_bwd = bufs
                w64[:self.n_out, :self.c_mid].copy_(hw_)
                w_fwd.copy_(w64.permute(0, 2, 3, 1).reshape(64, 9, 64))                  # [co, tap, ci]
                w_bwd.copy_(w64.flip(2, 3).permute(1, 2, 3, 0).reshape(64, 9, 64))       # [ci, flipped tap, co]
                out = head_conv_tc(m, w_fwd, hb, self.n_out)
            else:
                out = torch.empty((B, self.n_out, H, W), dtype=torch.float32, device=dev)
                # m has a 64-channel row stride (upper 32 are zero padding); the head reads only the 32 real channels
                _lib.check(lib.im2im_head_conv3x3_f32(m.data_ptr(), hw_.data_ptr(), hb.data_ptr(), None, B, H, W, self.c_mid,
                                                      64, self.n_out, out.data_ptr(), _st(dev)), "head_conv")
            ctx["head_w"] = hw_
            if self._tracked:
                torch._foreach_add_(self._tracked, 1)
            self._tracked = None
            self._zero = None
            if self.head_act is not None:
                # gaussian: relu(variance conv), residual magnitude: |magnitude conv| (gaussian_layer.py:18,
                # residual_magnitude_layer.py:18); the derivative (0/1 or sign) of the pre-activation is kept for backward
                tail = out[:, self.act_from:]
                ctx["act_grad"] = (tail > 0).float() if self.head_act == "relu" else torch.sign(tail)
                tail.copy_(torch.relu(tail) if self.head_act == "relu" else tail.abs())
        return out.view(B, self.n_planes, self.c_head, H, W), ctx

    # ------------------------------------------------------------------------------------------- backward
    def _conv_bwd(self, layer: _ConvBN, saved: dict, dz: torch.Tensor, grads: Dict, need_dx: bool = True,
                  bn_next=None):
        """wgrad (+ dgrad) of a tensor-core conv layer; returns (dx for x_in, dx for x_in2 or None).

        ``bn_next`` = (layer, saved) of the BatchNorm+ReLU layer whose output is this conv's (single) input: its backward
        is then folded in - the data gradient's epilogue applies the ReLU mask and accumulates the two per-channel sums
        (im2im_conv_igemm_bf16_stats, stat_mode 2), one apply pass finishes - and the first return value is that layer's dz.

        The weight gradient is off the critical path (nothing downstream in the backward pass reads it), so it is
        enqueued on a side stream: the tensor-bound wgrad GEMMs overlap the bandwidth-bound BatchNorm/ReLU backward
        kernels of the next layers running on the main stream."""
        x1, x2 = saved["x_in"], saved.get("x_in2")
        main = torch.cuda.current_stream(dz.device)
        side = self._side_stream(dz.device)
        c_out = dz.shape[3]
        # 1. the data gradient first, on the main stream: it is the critical path of the backward pass
        dx1 = dx2 = None
        fused = False
        if need_dx:
            if x2 is None and bn_next is not None:
                la, sa = bn_next
                sums = self._zero.take(2 * la.c_out)
                # Measured on B200 (tools/halo_modes_bench.py, batch 78): the backward fusion costs the halo kernel more
                # (the epilogue's bn_z reads: +0.38 ms on 64->64 @320^2) than the separate reduction pass it replaces, so it
                # is off by default; the forward fusion (stat_mode 1) is +6 % on the convolution and replaces a pass over z.
                if getattr(self, "fuse_bwd_stats", os.environ.get("IM2IM_FUSE_BWD_STATS") == "1"):
                    bn = la.bn
                    dx1, fused = conv_igemm_stats(dz, layer.w_bwd, 2, sums,
                                                  bn=(sa["z"], bn.weight.detach(), bn.bias.detach(), sa["mean"], sa["rstd"]))
                else:
                    dx1 = conv_igemm(dz, layer.w_bwd)
            elif x2 is None:
                dx1 = conv_igemm(dz, layer.w_bwd)
            else:
                c1 = x1.shape[3]
                dx1, dx2 = conv_igemm(dz, layer.w_bwd[:c1]), conv_igemm(dz, layer.w_bwd[c1:])
        # 2. the weight gradient on the side stream, ordered AFTER the data gradient (both are tensor-bound kernels that fill
        #    the GPU, so they cannot share it anyway): it then runs next to the bandwidth-bound BatchNorm / pool / upsample
        #    backward kernels that follow on the main stream, instead of delaying the data gradient they are waiting for
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dw1 = conv_wgrad(x1, dz, 9, out=self._zero.take(c_out, 9, x1.shape[3]))
            if x2 is None:
                grads[layer.conv.weight] = self._w_to_torch(dw1, layer.c_in)
            else:
                dw2 = conv_wgrad(x2, dz, 9, out=self._zero.take(c_out, 9, x2.shape[3]))
                c1, c2 = x1.shape[3], x2.shape[3]
                grads[layer.conv.weight] = torch.cat([self._w_to_torch(dw1, c1), self._w_to_torch(dw2, c2)], dim=1)
        for t in (dz, x1, x2):                      # keep the allocator from recycling them under the side stream
            if t is not None:
                t.record_stream(side)
        if not need_dx:
            return None, None
        # 3. BatchNorm + ReLU backward of the layer that produced this conv's input
        if x2 is None and bn_next is not None:
            la, sa = bn_next
            if not fused:
                return self._bn_relu_bwd(dx1, la, sa, grads), None
            bn, C, z = la.bn, la.c_out, sa["z"]
            dza = torch.empty_like(z)
            n_pix = z.shape[0] * z.shape[1] * z.shape[2]
            _lib.check(self.lib.im2im_bn_relu_bwd_apply_bf16(dx1.data_ptr(), z.data_ptr(), bn.weight.data_ptr(),
                                                             bn.bias.data_ptr(), sa["mean"].data_ptr(), sa["rstd"].data_ptr(),
                                                             sums.data_ptr(), n_pix, C, 1, dza.data_ptr(), _st(dz.device)),
                       "bn_relu_bwd_apply")
            grads[bn.bias] = sums[:C]
            grads[bn.weight] = sums[C:]
            return dza, None
        return dx1, dx2

    def backward(self, ctx: dict, dout: torch.Tensor) -> Dict[torch.Tensor, torch.Tensor]:
        lib, dev = self.lib, dout.device
        grads: Dict[torch.Tensor, torch.Tensor] = {}
        x = ctx["x"]
        B, c_in, H, W = x.shape
        dout = dout.contiguous().float().view(B, self.n_out, H, W)
        self._zero = _ZeroPool(self._bwd_zero_floats, dev)   # ONE fill for every accumulated-into gradient buffer
        self._zero.buf.record_stream(self._side_stream(dev))  # its slices are written by the side-stream weight gradients
        if self.head_act is not None:
            dout = dout.clone()
            dout[:, self.act_from:] *= ctx["act_grad"]
        with torch.cuda.device(dev):
            m, y_last = ctx["m"], ctx["y_last"]
            if ctx.get("head_tc"):
                # head gradients on tensor cores: dOut as a 64-channel NHWC bf16 operand, weight gradient through the halo
                # wgrad kernel, data gradient = the halo conv with the flipped / transposed head weights
                buf = self.__dict__.get("_dout_nhwc")      # channels >= n_out stay zero for the engine's lifetime
                if buf is None or tuple(buf.shape) != (B, H, W, 64) or buf.device != dev:
                    buf = torch.zeros((B, H, W, 64), dtype=torch.bfloat16, device=dev)
                    self._dout_nhwc = buf
                dout_nhwc = planar_to_nhwc64(dout, out=buf if self.n_out <= 8 else None)
                dm = conv_igemm(dout_nhwc, self._head_bufs[2])                               # channels >= c_mid come out zero
                side = self._side_stream(dev)                # weight gradients after the data gradient, off the critical path
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    dw64 = conv_wgrad(m, dout_nhwc, 9, out=self._zero.take(64, 9, 64))      # [plane, tap, feature]
                dwh = dw64[:self.n_out, :, :self.c_mid].permute(0, 2, 1).reshape(self.n_out, self.c_mid, 3, 3)
                dbh = dout.sum(dim=(0, 2, 3))
            else:
                dm = torch.empty_like(m)
                dwh = self._zero.take(self.n_out, self.c_mid, 3, 3)
                dbh = self._zero.take(self.n_out)
                _lib.check(lib.im2im_head_bwd(dout.data_ptr(), m.data_ptr(), ctx["head_w"].data_ptr(), B, H, W, self.c_mid,
                                              64, self.n_out, dm.data_ptr(), dwh.data_ptr(), dbh.data_ptr(), _st(dev)),
                           "head_bwd")
            co = self.c_head
            for i, conv in enumerate(self.head_convs):
                grads[conv.weight] = dwh[i * co:(i + 1) * co]
                grads[conv.bias] = dbh[i * co:(i + 1) * co]
            # 1x1 out conv: dgrad, then (side stream) wgrad; bias grad = channel sums of dm
            w_out_bwd = ctx["w_out_pad"].permute(2, 1, 0).contiguous()  # [64 ci, 1, 64 co]
            dy = conv_igemm(dm, w_out_bwd)
            side = self._side_stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                dw_out = conv_wgrad(y_last, dm, 1, out=self._zero.take(64, 1, 64))   # [64 (32 real), 1, 64]
            dm.record_stream(side)
            grads[self.out_conv.weight] = dw_out[:self.c_mid].view(self.c_mid, -1, 1, 1)
            sums = self._zero.take(128)
            _lib.check(lib.im2im_channel_stats_bf16(dm.data_ptr(), B * H * W, 64, sums.data_ptr(), _st(dev)), "dbias")
            grads[self.out_conv.bias] = sums[:self.c_mid]
            # up path, reversed
            skip_grads: List[Optional[torch.Tensor]] = [None] * 4      # for x1..x4
            for k in range(3, -1, -1):
                (a, b), (sa, sb) = self.up[k], ctx["up"][k]
                dz = self._bn_relu_bwd(dy, b, sb, grads)
                dz, _ = self._conv_bwd(b, sb, dz, grads, bn_next=(a, sa))     # dz of layer a
                d_skip, d_u = self._conv_bwd(a, sa, dz, grads)
                skip_grads[3 - k] = d_skip                              # up1 uses x4, ..., up4 uses x1
                lb, lh, lw, lc = sa["low_shape"]
                dy = torch.empty(sa["low_shape"], dtype=torch.bfloat16, device=dev)
                _lib.check(lib.im2im_upsample2x_bilinear_bwd_bf16(d_u.data_ptr(), lb, lh, lw, lc, d_u.shape[1],
                                                                  d_u.shape[2], dy.data_ptr(), _st(dev)), "upsample_bwd")
            # dy is now the gradient of x5; down path, reversed
            skips = ctx["skips"]                                        # [x1, x2, x3, x4, x5]
            d_p = None                                                  # gradient of the pooled tensor below a skip layer
            for k in range(3, -1, -1):
                (a, b), (sa, sb) = self.down[k], ctx["down"][k]
                if d_p is None:
                    dz = self._bn_relu_bwd(dy, b, sb, grads)            # x5: no pool below it
                else:                                                   # x2..x4: skip gradient + the pool's, one pass
                    dz = self._bn_relu_pool_bwd(skip_grads[k + 1], d_p, b, sb, grads)
                dz, _ = self._conv_bwd(b, sb, dz, grads, bn_next=(a, sa))     # dz of layer a
                d_p, _ = self._conv_bwd(a, sa, dz, grads)               # gradient of this block's max-pool output
            # inc block: x1 = skip gradient from up4 + the gradient through down1's pool
            s0, s1 = ctx["inc"]
            dz = self._bn_relu_pool_bwd(skip_grads[0], d_p, self.inc[1], s1, grads)
            dz0, _ = self._conv_bwd(self.inc[1], s1, dz, grads, bn_next=(self.inc[0], s0))
            first = self.inc[0]
            dw0 = self._zero.take(first.c_out, c_in, 3, 3)
            _lib.check(lib.im2im_conv_first_wgrad(x.data_ptr(), dz0.data_ptr(), B, c_in, H, W, first.c_out,
                                                  dw0.data_ptr(), _st(dev)), "conv_first_wgrad")
            grads[first.conv.weight] = dw0
            torch.cuda.current_stream(dev).wait_stream(self._side_stream(dev))  # weight gradients are complete
        return grads


class _NativeTrainFn(torch.autograd.Function):
    """Autograd bridge: forward/backward of the whole UNet + head run on the native engine; parameter gradients are
    returned to autograd so ``loss.backward()`` / any torch optimizer work as in the reference's loop."""

    @staticmethod
    def forward(ctx, x, engine, *params):
        out, saved = engine.forward(x)
        ctx.engine, ctx.saved, ctx.params = engine, saved, params
        return out

    @staticmethod
    def backward(ctx, dout):
        grads = ctx.engine.backward(ctx.saved, dout)
        ctx.saved = None
        return (None, None) + tuple(grads.get(p) for p in ctx.params)


def native_train_applicable(model, x) -> bool:
    from .unet import UNet
    from .unet_engine import head_plane_convs
    if not (model.training and torch.is_tensor(x) and x.is_cuda and torch.is_grad_enabled()
            and type(model.baseModel) is UNet and model.baseModel.bilinear
            and getattr(model, "use_native_training", True)
            and x.dim() == 4 and x.shape[1] <= 8 and min(x.shape[2], x.shape[3]) >= 16):
        return False
    planes = head_plane_convs(model.last_layer)   # quantile / gaussian / residual / quantile-l1 / inn heads
    return planes is not None and len(planes[0]) * planes[0][0].weight.shape[0] in (2, 3, 4, 6)


def native_train_forward(model, x):
    eng = model.__dict__.get("_native_train_engine")
    if eng is None:
        eng = UNetTrainEngine(model)
        model.__dict__["_native_train_engine"] = eng
    params = tuple(model.parameters())
    return _NativeTrainFn.apply(x, eng, *params)


# ------------------------------------------------------------------------------------------------ loss + optimizer
class _QuantileLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, q_lo, q_hi, w_lo, w_hi, w_mse):
        lib = _lib.load()
        pred_c, target_c = pred.contiguous(), target.contiguous().float()
        n = pred_c.shape[0]
        px = target_c[0].numel()
        dpred = torch.empty_like(pred_c)
        parts = torch.empty(3, dtype=torch.float64, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(lib.im2im_quantile_loss_f32(pred_c.data_ptr(), target_c.data_ptr(), n, px, q_lo, q_hi, w_lo, w_hi,
                                                   w_mse, dpred.data_ptr(), parts.data_ptr(), _st(pred.device)),
                       "quantile_loss")
        ctx.save_for_backward(dpred)
        count = float(n * px)
        # python-scalar weights: no host->device copy, so the step can be captured into a CUDA graph
        return ((parts[0] * w_lo + parts[1] * w_hi + parts[2] * w_mse) / count).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g, None, None, None, None, None, None


def native_quantile_loss(pred, target, params):
    """quantile_regression_loss_fn on the fused kernel (pred (B,3,C,H,W) fp32 CUDA, target (B,C,H,W))."""
    return _QuantileLossFn.apply(pred, target, float(params["q_lo"]), float(params["q_hi"]),
                                 float(params["q_lo_weight"]), float(params["q_hi_weight"]), float(params["mse_weight"]))


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (defaults of core/scripts/train.py:120) with ONE kernel over a flat fp32 buffer.

    Parameters are re-pointed at views of one contiguous buffer (as DDP-style flat buckets), gradients likewise, so the
    data-parallel all-reduce is a single NCCL call on ``flat_grad`` and the update a single launch."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        params = [p for p in params]
        if weight_decay != 0.0 or amsgrad:
            raise ValueError("FusedAdam implements torch.optim.Adam's defaults only (no weight decay, no amsgrad) - the "
                             "configuration of core/scripts/train.py:120")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam updates ONE flat buffer with one (lr, betas, eps): pass a single parameter group "
                             f"(got {len(self.param_groups)})")
        ps = [p for g in self.param_groups for p in g["params"]]
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        off = 0
        for p in ps:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p.data)
            p.grad = self.flat_grad[off:off + k].view_as(p.data)
            off += k
        self._params = ps
        self._step = 0
        # {step, 1-b1^step, sqrt(1-b2^step)} in device memory: the step count advances on the device, so a captured
        # CUDA graph of the whole training step replays with the right bias correction
        self.state_dev = torch.zeros(3, dtype=torch.float32, device=dev)

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grad.zero_()
        off = 0
        for p in self._params:  # autograd may have replaced .grad with a fresh tensor; point it back at the flat view
            k = p.numel()
            p.grad = self.flat_grad[off:off + k].view_as(p.data)
            off += k

    def gather_grads(self):
        """Copy gradients that autograd left outside the flat buffer into it (no-op when they already are views)."""
        off = 0
        for p in self._params:
            k = p.numel()
            view = self.flat_grad[off:off + k]
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad.reshape(-1))
            off += k

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        self.gather_grads()
        self._step += 1
        _lib.weights_changed()          # parameters are rewritten through raw pointers: tensor._version does not move
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam: parameter groups were added after construction; one group is supported")
        g = self.param_groups[0]
        lib = _lib.load()
        with torch.cuda.device(self.flat_param.device):
            _lib.check(lib.im2im_adam_step_dev_f32(self.flat_param.data_ptr(), self.flat_grad.data_ptr(),
                                                   self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                                   self.flat_param.numel(), g["lr"], g["betas"][0], g["betas"][1],
                                                   g["eps"], self.state_dev.data_ptr(), grad_scale,
                                                   _st(self.flat_param.device)), "adam_step")


class GraphedTrainStep:
    """The reference's training iteration (core/scripts/train.py:152-162: forward, loss_fn, zero_grad, backward,
    optimizer.step) captured ONCE into a CUDA graph and replayed - ~400 kernel launches per step become one
    cudaGraphLaunch, which removes the host launch overhead that otherwise bounds the step.  Data parallel: pass a
    torch.distributed ``group``; the NCCL all-reduce of the flat gradient buffer is part of the graph.

        step = GraphedTrainStep(model, FusedAdam(model.parameters(), lr=1e-4), x0, y0)
        for x, y in loader: loss = step(x, y)          # 0-dim CUDA tensor; .item() it when you need the number

    Shapes are fixed at capture (the reference's loader uses a fixed batch size; a ragged last batch should run the
    eager path).  BatchNorm running statistics, ``num_batches_tracked`` and Adam's step counter all advance on the device.
    """

    def __init__(self, model, optimizer: "FusedAdam", x: torch.Tensor, y: torch.Tensor, group=None, warmup: int = 3,
                 keep_warmup: bool = False):
        assert x.is_cuda and y.is_cuda and model.training
        self.model, self.opt, self.group = model, optimizer, group
        self.world = 1
        if group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
        self.x, self.y = x.clone(), y.clone()
        dev = x.device
        # The warm-up iterations (lazy allocations, NCCL channel setup) are real optimizer steps on the example batch.  Unless
        # the caller asks to keep them, parameters, Adam moments / step counter and BatchNorm buffers are put back afterwards,
        # so that the first replay is the FIRST step of the trajectory, exactly as in the reference's loop (train.py:147-162).
        snapshot = None
        if not keep_warmup:
            snapshot = ([t.clone() for t in (optimizer.flat_param, optimizer.exp_avg, optimizer.exp_avg_sq,
                                             optimizer.state_dev)],
                        [b.clone() for b in model.buffers() if b is not None], optimizer._step)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):            # warm-up off the capture: lazy allocations, cuDNN-free, NCCL channels
            for _ in range(warmup):
                self._iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if snapshot is not None:
            with torch.no_grad():
                for dst, src in zip((optimizer.flat_param, optimizer.exp_avg, optimizer.exp_avg_sq, optimizer.state_dev),
                                    snapshot[0]):
                    dst.copy_(src)
                for dst, src in zip([b for b in model.buffers() if b is not None], snapshot[1]):
                    dst.copy_(src)
            optimizer._step = snapshot[2]
            _lib.weights_changed()
            torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = self._iteration()
        self.kernels_per_replay = _lib.launch_count() - before
        # optimizer steps already taken on the example batch when the first replay runs (capture itself executes nothing)
        self.warmup_steps = warmup if keep_warmup else 0

    def _iteration(self):
        self.opt.zero_grad()
        loss = self.model.loss_fn(self.model(self.x), self.y)
        loss.backward()
        if self.group is not None:
            import torch.distributed as dist
            self.opt.gather_grads()
            dist.all_reduce(self.opt.flat_grad, group=self.group)
        self.opt.step(grad_scale=1.0 / self.world)
        return loss.detach()

    def close(self):
        """Release the captured graph and its memory pool (do this before destroying a process group whose collectives
        were captured: NCCL waits for such graphs at communicator teardown)."""
        torch.cuda.synchronize(self.x.device)
        if self.graph is not None:
            self.graph.reset()
            self.graph = None
        self.loss = None

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        _lib.weights_changed()          # the replay rewrites parameters and BatchNorm statistics behind torch's back
        self.graph.replay()
        return self.loss
