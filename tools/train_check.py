"""Bring-up check of the native training step against torch autograd through the same modules (fp32, cuDNN)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from core.models.add_uncertainty import add_uncertainty
from core.models.trunks.unet import UNet
from im2im_uq_b200.models.unet_train import FusedAdam

torch.backends.cudnn.allow_tf32 = False
params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
dev = "cuda:0"

def build(seed=0):
    torch.manual_seed(seed)
    return add_uncertainty(UNet(1, 1), params).to(dev).train()

def run(model, x, y, native):
    model.use_native_training = native
    model.zero_grad(set_to_none=True)
    pred = model(x)
    loss = model.loss_fn(pred, y)
    loss.backward()
    return pred.detach(), loss.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(B, 1, H, W, device=dev, generator=g)
y = x + 0.3 * torch.randn(B, 1, H, W, device=dev, generator=g)
m_ref, m_nat = build(), build()
p_ref, l_ref, g_ref = run(m_ref, x, y, False)
p_nat, l_nat, g_nat = run(m_nat, x, y, True)
torch.cuda.synchronize()
rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
# how far does torch's own bf16 autocast land from its fp32 result?  (the yardstick for "bf16-level agreement")
m_ac = build()
m_ac.use_native_training = False
with torch.autocast("cuda", dtype=torch.bfloat16):
    pred_ac = m_ac(x)
loss_ac = m_ac.loss_fn(pred_ac.float(), y)
loss_ac.backward()
g_ac = {n: p.grad.detach().clone() for n, p in m_ac.named_parameters()}
ac = sorted(rel(g_ac[n], g_ref[n]) for n in g_ref if g_ref[n].norm() > 1e-7)
print(f"torch autocast-bf16 vs torch fp32: pred rel L2 {rel(pred_ac.float(), p_ref):.3e}; grad rel err median {ac[len(ac)//2]:.3e} max {ac[-1]:.3e}")
print(f"pred rel L2 {rel(p_nat, p_ref):.3e}   loss native {l_nat.item():.6f} ref {l_ref.item():.6f}")
worst = []
for n in g_ref:
    r = rel(g_nat[n], g_ref[n]) if g_ref[n].norm() > 1e-7 else float(g_nat[n].norm())
    worst.append((r, n, g_ref[n].norm().item()))
worst.sort(reverse=True)
for r, n, nr in worst[:12]: print(f"  grad rel err {r:.3e}  |ref| {nr:.3e}  {n}")
med = float(np.median([w[0] for w in worst]))
print(f"median grad rel err {med:.3e}; max {worst[0][0]:.3e}")
# running stats parity
for (n1, b1), (n2, b2) in zip(m_ref.named_buffers(), m_nat.named_buffers()):
    if b1 is not None and b1.dtype.is_floating_point and rel(b2.float(), b1.float()) > 2e-2:
        print("  running stat mismatch", n1, rel(b2.float(), b1.float()))
# a few Adam steps: loss should go down identically-ish
opt_ref = torch.optim.Adam(m_ref.parameters(), lr=1e-3)
opt_nat = FusedAdam(m_nat.parameters(), lr=1e-3)
for it in range(5):
    for model, opt, native in ((m_ref, opt_ref, False), (m_nat, opt_nat, True)):
        model.use_native_training = native
        opt.zero_grad()
        loss = model.loss_fn(model(x), y)
        loss.backward()
        opt.step()
        print(f"  it {it} {'native' if native else 'torch '} loss {loss.item():.6f}", end="")
    print()
# speed at 320x320
B2 = 16
x2 = torch.randn(B2, 1, 320, 320, device=dev); y2 = x2 + 0.1
def timeit(model, opt, native, iters=5):
    model.use_native_training = native
    def step():
        opt.zero_grad(); loss = model.loss_fn(model(x2), y2); loss.backward(); opt.step(); return loss
    for _ in range(2): step()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(iters): l = step()
    l.item(); torch.cuda.synchronize()
    return (time.perf_counter() - t) / iters
t_nat = timeit(m_nat, opt_nat, True)
print(f"native train step B={B2}: {t_nat * 1e3:.1f} ms  {B2 / t_nat:.0f} img/s")
torch.backends.cudnn.allow_tf32 = True
t_ref = timeit(m_ref, opt_ref, False)
print(f"torch (tf32 cuDNN) train step B={B2}: {t_ref * 1e3:.1f} ms  {B2 / t_ref:.0f} img/s")
