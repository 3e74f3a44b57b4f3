"""ncu target: the CTA-pair halo convolution at batch 78 on the UNet's wide layers, plus one halo weight gradient.
ncu --set full --clock-control none --import-source on -k regex:"conv_halo|conv_wgrad_halo" -c 6 python tools/pair_profile.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200 import conv
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 78
for (H, c1, c2, co, stats) in [(320, 64, 0, 64, False), (320, 64, 0, 64, True), (320, 64, 64, 64, False), (160, 128, 0, 128, False),
                               (160, 64, 0, 128, False)]:
    x1 = torch.randn(B, H, H, c1, device=dev).to(torch.bfloat16)
    x2 = torch.randn(B, H, H, c2, device=dev).to(torch.bfloat16) if c2 else None
    w = (torch.randn(co, 9, c1 + c2, device=dev) / 30).to(torch.bfloat16)
    b = torch.randn(co, device=dev)
    if stats:
        sums = torch.zeros(2 * co, device=dev)
        conv.conv_igemm_stats(x1, w, 1, sums, x2=x2)
    else:
        conv.conv_igemm(x1, w, b, True, x2)
    torch.cuda.synchronize()
x = torch.randn(B, 320, 320, 64, device=dev).to(torch.bfloat16)
dz = torch.randn(B, 320, 320, 64, device=dev).to(torch.bfloat16)
conv.conv_wgrad(x, dz, 9)
torch.cuda.synchronize()
print("done")
