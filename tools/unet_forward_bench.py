"""UNet + quantile head inference forward: native engine vs torch (cuDNN) bf16/tf32, images/s at 320x320."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from core.models.add_uncertainty import add_uncertainty
from core.models.trunks.unet import UNet
from im2im_uq_b200 import _lib

params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
torch.manual_seed(0)
model = add_uncertainty(UNet(1, 1), params).to("cuda:0").eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = torch.randn(B, 1, 320, 320, device="cuda:0")
FL = 125.29e9

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

with torch.no_grad():
    l0 = _lib.launch_count()
    model(x); torch.cuda.synchronize()
    print("launches per forward:", _lib.launch_count() - l0)
    ms = timeit(lambda: model(x))
    print(f"native  B={B}: {ms:.2f} ms  {B / ms * 1e3:.0f} img/s  {B * FL / ms * 1e-9:.0f} TFLOP/s")
    model.use_native_inference = False
    ms = timeit(lambda: model(x))
    print(f"torch fp32(tf32 conv) B={B}: {ms:.2f} ms  {B / ms * 1e3:.0f} img/s")
    mb = model.to(torch.bfloat16).to(memory_format=torch.channels_last)
    xb = x.to(torch.bfloat16).to(memory_format=torch.channels_last)
    ms = timeit(lambda: mb(xb))
    print(f"torch bf16 channels_last B={B}: {ms:.2f} ms  {B / ms * 1e3:.0f} img/s")
