#!/usr/bin/env python
"""Per-kernel GPU time of the native training step / inference forward (CUPTI via torch.profiler; kernels launched
through ctypes are captured too).  usage: train_profile.py [train|fwd] [batch] [side]"""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from core.models.add_uncertainty import add_uncertainty
from core.models.trunks.unet import UNet
from im2im_uq_b200.models.unet_train import FusedAdam

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 320
dev = "cuda:0"
params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
torch.manual_seed(0)
model = add_uncertainty(UNet(1, 1), params).to(dev)
x = torch.randn(B, 1, S, S, device=dev)
y = x + 0.3 * torch.randn_like(x)
if mode == "train":
    model.train()
    opt = FusedAdam(model.parameters(), lr=1e-4)
    def step():
        opt.zero_grad()
        loss = model.loss_fn(model(x), y)
        loss.backward()
        opt.step()
        return loss
else:
    model.eval()
    def step():
        with torch.no_grad():
            return model(x)
for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
print(f"{mode} B={B} {S}x{S}: {e0.elapsed_time(e1) / 10:.3f} ms/step (events, 10 steps)")
N = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("void ", "").replace("im2im::(anonymous namespace)::", "")[:90]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
if os.environ.get("IM2IM_PROFILE_LIST"):      # every launch of the last profiled step, in launch order
    evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    per = len(evs) // N
    for ev in evs[-per:]:
        d = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        print(f"  {d:9.1f} us  {ev.name.replace('void ', '').replace('im2im::(anonymous namespace)::', '')[:60]}")
tot = sum(v[1] for v in agg.values())
print(f"sum of kernel time {tot / N / 1e3:.3f} ms/step over {sum(v[0] for v in agg.values()) // N} launches/step")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1] / N / 1e3:9.3f} ms {100 * v[1] / tot:6.2f}%  n={v[0] // N:4d}  avg {v[1] / v[0]:9.1f} us  {k}")
