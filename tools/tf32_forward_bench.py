"""UNet + quantile head inference forward in both native precisions (bf16 and the reference's, kind::tf32), batch 78, 320x320;
IM2IM_TF32_HALO=0 keeps the tf32 mode's wide layers on the persistent kernel (A/B)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from core.models.add_uncertainty import add_uncertainty
from core.models.trunks.unet import UNet

params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
torch.manual_seed(0)
model = add_uncertainty(UNet(1, 1), params).to("cuda:0").eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 78
x = torch.randn(B, 1, 320, 320, device="cuda:0")


def timeit(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


with torch.no_grad():
    for prec, halo in (("bf16", "1"), ("tf32", "1"), ("tf32", "0"), ("tf32", "1"), ("bf16", "1")):
        os.environ["IM2IM_TF32_HALO"] = halo
        model.native_precision = prec
        ms = timeit(lambda: model(x))
        print(f"{prec} (IM2IM_TF32_HALO={halo}) B={B}: {ms:.2f} ms  {B / ms * 1e3:.0f} images/s", flush=True)
