"""ncu target: three native training steps at batch 78 (no torch.profiler - CUPTI has one subscriber).
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:conv_ --launch-skip 130 --launch-count 65 --csv python tools/train_step_ncu_target.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from core.models.add_uncertainty import add_uncertainty
from core.models.trunks.unet import UNet
from im2im_uq_b200.models.unet_train import FusedAdam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 78
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
dev = "cuda:0"
params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
torch.manual_seed(0)
model = add_uncertainty(UNet(1, 1), params).to(dev)
x = torch.randn(B, 1, 320, 320, device=dev)
y = x + 0.3 * torch.randn_like(x)
if mode == "train":
    model.train()
    opt = FusedAdam(model.parameters(), lr=1e-4)
    for _ in range(3):
        opt.zero_grad()
        loss = model.loss_fn(model(x), y)
        loss.backward()
        opt.step()
else:
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            model(x)
torch.cuda.synchronize()
print("done")
