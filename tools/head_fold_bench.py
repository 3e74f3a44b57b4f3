import os, sys
sys.path.insert(0, "/root/repo")
import torch
from im2im_uq_b200.conv import (conv_igemm, fold_outconv_into_head, head_conv_tc, pack_conv_weight, pad_head_weight)
dev = "cuda:0"
B, S = 78, 320
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
g = torch.Generator().manual_seed(0)
y = torch.relu(torch.randn(B, S, S, 64, device=dev)).to(torch.bfloat16)
ow, ob = (torch.randn(32, 64, 1, 1, generator=g) * 0.1).to(dev), torch.randn(32, generator=g).to(dev)
hw, hb = (torch.randn(3, 32, 3, 3, generator=g) * 0.1).to(dev), torch.randn(3, generator=g).to(dev)
ow64 = torch.zeros(64, 64, 1, 1, device=dev); ow64[:32] = ow
ob64 = torch.zeros(64, device=dev); ob64[:32] = ob
out64 = pack_conv_weight(ow64)
head_p = pack_conv_weight(pad_head_weight(hw))
wf, bf, tb = fold_outconv_into_head(hw, hb, ow, ob)
fold_p = pack_conv_weight(pad_head_weight(wf))
m = conv_igemm(y, out64, ob64, relu=False)
print("outc            %.3f ms" % t(lambda: conv_igemm(y, out64, ob64, relu=False)))
print("head on m       %.3f ms" % t(lambda: head_conv_tc(m, head_p, hb, 3)))
print("folded on y     %.3f ms" % t(lambda: head_conv_tc(y, fold_p, bf, 3, tap_bias=tb)))
print("folded, no tapb %.3f ms" % t(lambda: head_conv_tc(y, fold_p, bf, 3)))
print("head_p on y     %.3f ms" % t(lambda: head_conv_tc(y, head_p, hb, 3)))
print("fold_p on m     %.3f ms" % t(lambda: head_conv_tc(m, fold_p, bf, 3)))
