#!/usr/bin/env python
"""One-pass RCPS kernel per head kind: achieved HBM GB/s (algorithmic bytes / CUDA-event time) at 4k x 320^2, L = 1000."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200 import _lib, rcps

dev = torch.device("cuda:0")
n, h, w, L = 4000, 320, 320, 1000
g = torch.Generator(device=dev).manual_seed(0)
shape = (n, 1, h, w)
pred = torch.rand(shape, generator=g, device=dev)
sig = 0.02 + 0.1 * torch.rand(shape, generator=g, device=dev)
lab = pred + sig * torch.randn(shape, generator=g, device=dev)
width = sig * (0.5 + torch.rand(shape, generator=g, device=dev))
lam = torch.linspace(0.0, 6.0, L, device=dev)
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6538.9
cases = {
    "QUANTILES (16 B/px)": (_lib.IM2IM_HEAD_QUANTILES, lambda: torch.stack([pred - width, pred, pred + width], 1).contiguous(), 16),
    "RESIDUAL (12 B/px)": (_lib.IM2IM_HEAD_RESIDUAL, lambda: torch.stack([pred, width], 1).contiguous(), 12),
    "GAUSSIAN (12 B/px)": (_lib.IM2IM_HEAD_GAUSSIAN, lambda: torch.stack([pred, width ** 2], 1).contiguous(), 12),
    "SOFTMAX_SETS (16 B/px)": (_lib.IM2IM_HEAD_SOFTMAX_SETS, lambda: torch.stack(
        [torch.floor((pred - width).clamp(0, 1) * 50) / 50, torch.floor(pred * 50) / 50,
         torch.floor((pred + width).clamp(0, 1) * 50) / 50], 1).contiguous(), 16),
}
counts = torch.empty((n, L), dtype=torch.int32, device=dev)
totals = torch.empty((L,), dtype=torch.int64, device=dev)
for name, (kind, make, bpp) in cases.items():
    out = make()
    for _ in range(3):
        rcps.miss_counts(out, lab, lam, counts=counts, totals=totals, zero=True, head=kind)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    a.record()
    for _ in range(iters):
        rcps.miss_counts(out, lab, lam, counts=counts, totals=totals, zero=False, head=kind)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    alg = n * h * w * bpp + n * L * 4
    print(f"{name:24s} {ms:7.3f} ms  {alg / ms / 1e6:8.1f} GB/s  = {alg / ms / 1e6 / peak:5.3f} of the measured copy peak "
          f"({n / ms * 1e3 / 1e6:.2f} M images/s)", flush=True)
    del out
# softmax set extraction: logits (n2, 50, 1, h, w) read once, 3 planes written
n2 = 400
logits = torch.randn(n2, 50, 1, h, w, device=dev)
for _ in range(2):
    rcps.softmax_sets(logits)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    rcps.softmax_sets(logits)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
alg = n2 * h * w * (50 * 4 + 12)
print(f"softmax_sets (K=50, 212 B/px) {ms:7.3f} ms  {alg / ms / 1e6:8.1f} GB/s = {alg / ms / 1e6 / peak:5.3f} of peak "
      f"({n2 / ms * 1e3:.0f} images/s)")
