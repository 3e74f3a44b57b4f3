"""Run a few native UNet train steps (for ncu launch lists / timing). usage: train_profile.py [batch] [side] [steps]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from core.models.add_uncertainty import add_uncertainty
from core.models.trunks.unet import UNet
from im2im_uq_b200.models.unet_train import FusedAdam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
side = int(sys.argv[2]) if len(sys.argv) > 2 else 320
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
torch.manual_seed(0)
model = add_uncertainty(UNet(1, 1), params).to("cuda:0").train()
opt = FusedAdam(model.parameters(), lr=1e-4)
x = torch.randn(B, 1, side, side, device="cuda:0"); y = x + 0.1
def step():
    opt.zero_grad(); loss = model.loss_fn(model(x), y); loss.backward(); opt.step(); return loss
for _ in range(2): step()
torch.cuda.synchronize(); t = time.perf_counter()
per = []
for _ in range(steps):
    t1 = time.perf_counter(); l = step(); l.item(); per.append((time.perf_counter() - t1) * 1e3)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / steps
print("per-step ms:", " ".join(f"{p:.1f}" for p in per), "| max mem GB", round(torch.cuda.max_memory_allocated() / 1e9, 2),
      "reserved", round(torch.cuda.memory_reserved() / 1e9, 2))
print(f"native train step B={B} {side}x{side}: {dt * 1e3:.2f} ms  {B / dt:.0f} img/s  loss {l.item():.4f}")
