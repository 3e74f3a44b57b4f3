"""Bring-up check for the tcgen05 implicit-GEMM convolution: numerics vs torch fp32 conv on bf16-rounded operands, timing."""
import sys, os, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200 import conv

dev = torch.device("cuda:0")
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

def check(B, H, W, cin1, cin2, cout, taps, relu=True, bias=True, out_dtype=torch.bfloat16):
    k = 3 if taps == 9 else 1
    x1 = torch.randn(B, cin1, H, W, device=dev)
    x2 = torch.randn(B, cin2, H, W, device=dev) if cin2 else None
    w = torch.randn(cout, cin1 + cin2, k, k, device=dev) / ((cin1 + cin2) * taps) ** 0.5
    b = torch.randn(cout, device=dev) if bias else None
    x1b, wb = x1.to(torch.bfloat16), w.to(torch.bfloat16)
    x2b = x2.to(torch.bfloat16) if cin2 else None
    xin = x1b.float() if not cin2 else torch.cat([x1b.float(), x2b.float()], dim=1)
    ref = F.conv2d(xin, wb.float(), b, padding=k // 2)
    if relu: ref = ref.relu()
    got = conv.conv_igemm(conv.to_nhwc_bf16(x1b), conv.pack_conv_weight(wb), b, relu,
                          conv.to_nhwc_bf16(x2b) if cin2 else None, out_dtype)
    torch.cuda.synchronize()
    got = got.float().permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-2 * scale if out_dtype == torch.bfloat16 else 2e-3 * scale
    ok = err <= tol
    print(f"B{B} {H}x{W} cin {cin1}+{cin2} cout {cout} taps {taps} {str(out_dtype)[6:]}: max|err|={err:.4g} (scale {scale:.3g}) {'OK' if ok else 'FAIL'}", flush=True)
    return ok

ok = True
ok &= check(1, 16, 16, 64, 0, 64, 9)
ok &= check(2, 16, 8, 64, 0, 32, 1, relu=False, out_dtype=torch.float32)
ok &= check(2, 40, 40, 128, 0, 128, 9)
ok &= check(3, 20, 20, 256, 0, 512, 9)
ok &= check(2, 32, 32, 64, 64, 64, 9)
ok &= check(1, 37, 29, 64, 0, 256, 9, bias=False)
ok &= check(2, 64, 64, 128, 128, 64, 9, out_dtype=torch.float32)
print("NUMERICS", "OK" if ok else "FAIL", flush=True)

def bench(B, H, W, cin1, cin2, cout, taps=9, iters=20):
    x1 = torch.randn(B, H, W, cin1, device=dev).to(torch.bfloat16)
    x2 = torch.randn(B, H, W, cin2, device=dev).to(torch.bfloat16) if cin2 else None
    w = (torch.randn(cout, taps, cin1 + cin2, device=dev) / 30).to(torch.bfloat16)
    b = torch.randn(cout, device=dev)
    for _ in range(3): conv.conv_igemm(x1, w, b, True, x2)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): conv.conv_igemm(x1, w, b, True, x2)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / iters
    fl = 2.0 * B * H * W * cout * taps * (cin1 + cin2)
    # cuDNN bf16 NHWC reference
    xc = (x1 if not cin2 else torch.cat([x1, x2], dim=3)).permute(0, 3, 1, 2)  # channels_last view
    wc = w.view(cout, 3, 3, cin1 + cin2).permute(0, 3, 1, 2) if taps == 9 else w.view(cout, 1, 1, cin1 + cin2).permute(0, 3, 1, 2)
    bb = b.to(torch.bfloat16)
    for _ in range(3): F.conv2d(xc, wc, bb, padding=1 if taps == 9 else 0)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters): F.conv2d(xc, wc, bb, padding=1 if taps == 9 else 0)
    e.record(); torch.cuda.synchronize()
    ms2 = a.elapsed_time(e) / iters
    print(f"B{B} {H}x{W} {cin1}+{cin2}->{cout}: ours {ms:.3f} ms {fl / ms * 1e-9:.0f} TFLOP/s | cuDNN bf16 {ms2:.3f} ms {fl / ms2 * 1e-9:.0f} TFLOP/s", flush=True)

if ok:
    B = 16
    bench(B, 320, 320, 64, 0, 64)
    bench(B, 320, 320, 64, 64, 64)
    bench(B, 160, 160, 128, 0, 128)
    bench(B, 160, 160, 128, 128, 128)
    bench(B, 80, 80, 256, 0, 256)
    bench(B, 40, 40, 512, 0, 512)
    bench(B, 40, 40, 512, 512, 512)
    bench(B, 20, 20, 512, 0, 512)
