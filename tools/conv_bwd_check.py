"""Bring-up check: tcgen05 wgrad kernel and dgrad (igemm with flipped/transposed weights) vs torch autograd."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200 import conv

dev = torch.device("cuda:0")
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False

def check(B, H, W, cin, cout, taps=9):
    k = 3 if taps == 9 else 1
    x = torch.randn(B, cin, H, W, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, device=dev) / (cin * taps) ** 0.5).to(torch.bfloat16)
    dz = torch.randn(B, cout, H, W, device=dev).to(torch.bfloat16)
    xf = x.float().requires_grad_(True); wf = w.float().requires_grad_(True)
    y = F.conv2d(xf, wf, None, padding=k // 2)
    y.backward(dz.float())
    ref_dw = wf.grad.permute(0, 2, 3, 1).reshape(cout, taps, cin)       # [cout, tap, cin]
    ref_dx = xf.grad
    got_dw = conv.conv_wgrad(conv.to_nhwc_bf16(x), conv.to_nhwc_bf16(dz), taps)
    got_dx = conv.conv_igemm(conv.to_nhwc_bf16(dz), conv.pack_dgrad_weight(w), None, False, None, torch.float32)
    torch.cuda.synchronize()
    e_w = (got_dw - ref_dw).abs().max().item() / ref_dw.abs().max().item()
    e_x = (got_dx.permute(0, 3, 1, 2) - ref_dx).abs().max().item() / ref_dx.abs().max().item()
    ok = e_w < 2e-3 and e_x < 2e-3
    print(f"B{B} {H}x{W} {cin}->{cout} taps {taps}: wgrad rel err {e_w:.2e}  dgrad rel err {e_x:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok

ok = True
ok &= check(1, 16, 16, 64, 64)
ok &= check(2, 16, 8, 64, 128)
ok &= check(2, 40, 40, 128, 64)
ok &= check(3, 20, 20, 256, 512)
ok &= check(1, 37, 29, 64, 256)
ok &= check(2, 32, 32, 64, 64, taps=1)
ok &= check(2, 64, 64, 128, 128)
ok &= check(2, 48, 24, 128, 192)     # halo wgrad: several channel / c_out blocks, W not a multiple of 16
ok &= check(3, 32, 40, 64, 64)
print("BWD NUMERICS", "OK" if ok else "FAIL", flush=True)

def bench(B, H, W, cin, cout, iters=10):
    x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
    dz = torch.randn(B, H, W, cout, device=dev).to(torch.bfloat16)
    out = torch.zeros(cout, 9, cin, device=dev)
    for _ in range(2): conv.conv_wgrad(x, dz, 9, out)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): conv.conv_wgrad(x, dz, 9, out)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / iters
    fl = 2.0 * B * H * W * cout * 9 * cin
    print(f"wgrad B{B} {H}x{W} {cin}->{cout}: {ms:.3f} ms {fl / ms * 1e-9:.0f} TFLOP/s", flush=True)

if ok:
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    bench(B, 320, 320, 64, 64); bench(B, 320, 320, 128, 64); bench(B, 160, 160, 128, 128); bench(B, 80, 80, 256, 256)
    bench(B, 40, 40, 256, 512); bench(B, 40, 40, 512, 512); bench(B, 40, 40, 1024, 512); bench(B, 40, 40, 512, 256)
    bench(B, 20, 20, 512, 512)
