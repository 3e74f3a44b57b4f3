"""Dev bench: the halo convolution plain vs with fused statistics (forward sums / BatchNorm-backward sums)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from im2im_uq_b200.conv import conv_igemm, conv_igemm_stats

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 78


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for (H, c1, c2, co) in ((320, 64, 0, 64), (320, 64, 64, 64), (160, 64, 0, 128), (160, 128, 0, 128), (160, 128, 0, 64)):
    x1 = torch.randn(B, H, H, c1, device=dev).to(torch.bfloat16)
    x2 = torch.randn(B, H, H, c2, device=dev).to(torch.bfloat16) if c2 else None
    w = (torch.randn(co, 9, c1 + c2, device=dev) / 30).to(torch.bfloat16)
    z = torch.randn(B, H, H, co, device=dev).to(torch.bfloat16)
    gamma = torch.rand(co, device=dev) + 0.5; beta = torch.rand(co, device=dev); mean = torch.zeros(co, device=dev); rstd = torch.ones(co, device=dev)
    sums = torch.zeros(2 * co, device=dev)
    fl = 2.0 * B * H * H * co * 9 * (c1 + c2)
    ms0 = t(lambda: conv_igemm(x1, w, x2=x2))
    ms1 = t(lambda: conv_igemm_stats(x1, w, 1, sums, x2=x2))
    ms2 = t(lambda: conv_igemm_stats(x1, w, 2, sums, x2=x2, bn=(z, gamma, beta, mean, rstd)))
    print(f"B{B} {H}^2 {c1}+{c2}->{co}: plain {ms0:.3f} ms {fl/ms0*1e-9:.0f} TF | stats fwd {ms1:.3f} ms {fl/ms1*1e-9:.0f} TF | "
          f"stats bwd {ms2:.3f} ms {fl/ms2*1e-9:.0f} TF", flush=True)
