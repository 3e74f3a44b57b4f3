"""Dev bench: the halo weight gradient with N = 64 (IM2IM_WGRAD_HALO_WIDE=0) and N = 128 tiles at the UNet's wide layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from im2im_uq_b200 import conv
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 78
for (H, cin, cout) in ((320, 64, 64), (160, 64, 128), (160, 128, 128), (160, 256, 128), (80, 128, 256), (80, 256, 256), (80, 512, 256), (80, 256, 128)):
    x = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
    dz = torch.randn(B, H, H, cout, device=dev).to(torch.bfloat16)
    out = torch.zeros(cout, 9, cin, device=dev)
    res = []
    for wide in ("0", "1"):
        os.environ["IM2IM_WGRAD_HALO_WIDE"] = wide
        for _ in range(3): conv.conv_wgrad(x, dz, 9, out)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): conv.conv_wgrad(x, dz, 9, out)
        e.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(e) / 10)
    fl = 2.0 * B * H * H * cout * 9 * cin
    print(f"wgrad B{B} {H}^2 {cin}->{cout}: N=64 {res[0]:.3f} ms {fl / res[0] * 1e-9:.0f} TF | N=128 {res[1]:.3f} ms {fl / res[1] * 1e-9:.0f} TF", flush=True)
