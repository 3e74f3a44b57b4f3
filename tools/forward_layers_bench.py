"""Per-layer times of the inference forward's convolutions (BN folded, bias + ReLU, bf16 NHWC) at the reference's UNet shapes.
usage: python tools/forward_layers_bench.py [batch] [side]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from im2im_uq_b200.conv import conv_igemm  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 78
S = int(sys.argv[2]) if len(sys.argv) > 2 else 320


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


# (name, side divisor, c_in1 (skip), c_in2 (upsampled), c_out, taps)
LAYERS = [("inc.conv2", 1, 64, 0, 64, 9), ("down1.conv1", 2, 64, 0, 128, 9), ("down1.conv2", 2, 128, 0, 128, 9),
          ("down2.conv1", 4, 128, 0, 256, 9), ("down2.conv2", 4, 256, 0, 256, 9), ("down3.conv1", 8, 256, 0, 512, 9),
          ("down3.conv2", 8, 512, 0, 512, 9), ("down4.conv1", 16, 512, 0, 512, 9), ("down4.conv2", 16, 512, 0, 512, 9),
          ("up1.conv1", 8, 512, 512, 512, 9), ("up1.conv2", 8, 512, 0, 256, 9), ("up2.conv1", 4, 256, 256, 256, 9),
          ("up2.conv2", 4, 256, 0, 128, 9), ("up3.conv1", 2, 128, 128, 128, 9), ("up3.conv2", 2, 128, 0, 64, 9),
          ("up4.conv1", 1, 64, 64, 64, 9), ("up4.conv2", 1, 64, 0, 64, 9), ("outc (1x1, 64->64 padded)", 1, 64, 0, 64, 1)]
total_ms = total_fl = 0.0
for name, div, c1, c2, co, taps in LAYERS:
    H = S // div
    x1 = torch.randn(B, H, H, c1, device=dev).to(torch.bfloat16)
    x2 = torch.randn(B, H, H, c2, device=dev).to(torch.bfloat16) if c2 else None
    w = (torch.randn(co, taps, c1 + c2, device=dev) / 30).to(torch.bfloat16)
    bias = torch.randn(co, device=dev)
    fl = 2.0 * B * H * H * co * taps * (c1 + c2)
    ms = t(lambda: conv_igemm(x1, w, bias, relu=True, x2=x2))
    total_ms += ms
    total_fl += fl
    print(f"{name:28s} {H:3d}^2 {c1:4d}+{c2:<4d}->{co:4d}: {ms:7.3f} ms {fl / ms * 1e-9:6.0f} TFLOP/s", flush=True)
print(f"sum {total_ms:.3f} ms, {total_fl / total_ms * 1e-9:.0f} TFLOP/s over the convolutions (batch {B}, {S}x{S})")
