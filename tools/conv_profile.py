"""Launch a few forward / wgrad tensor-core convolutions of the UNet's shapes (target for `ncu --set full -k regex:conv_`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200 import conv
dev = torch.device("cuda:0")
B = 16
for (H, cin1, cin2, cout) in [(320, 64, 0, 64), (160, 128, 128, 128), (40, 512, 512, 512)]:
    x1 = torch.randn(B, H, H, cin1, device=dev).to(torch.bfloat16)
    x2 = torch.randn(B, H, H, cin2, device=dev).to(torch.bfloat16) if cin2 else None
    w = (torch.randn(cout, 9, cin1 + cin2, device=dev) / 30).to(torch.bfloat16)
    b = torch.randn(cout, device=dev)
    dz = torch.randn(B, H, H, cout, device=dev).to(torch.bfloat16)
    for _ in range(2):
        conv.conv_igemm(x1, w, b, True, x2)
        conv.conv_wgrad(x1, dz, 9)
torch.cuda.synchronize()
print("done")
