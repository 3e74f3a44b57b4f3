"""Dev probe: per-launch time of the unfused histogram kernel vs the fused calibration kernel, cold (after an idle pause)
and after a long back-to-back run (power-capped clocks), with NVML SM clocks sampled alongside."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
from bench import synth, CONFIGS
from im2im_uq_b200 import rcps
from im2im_uq_b200.calibration import calibrate_model as cm, sweep

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
clk = lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
pw = lambda: pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
out, lab = synth(n, 320, dev, 1000)
cfg = dict(alpha=0.1, delta=0.1, device="cuda:0", uncertainty_type="quantiles", minimum_lambda=0.0, maximum_lambda=6.0,
           num_lambdas=1000, rcps_loss="fraction_missed", dataset="synthetic")
lambdas, dlambda, lam_prime, _ = sweep.lambda_grid(cfg)
lam_dev = lam_prime.to(dev)
counts = torch.zeros((n, 1000), dtype=torch.int32, device=dev)
totals = torch.zeros(1000, dtype=torch.int64, device=dev)
plan_f = cm.RcpsGraph(out, lab, cfg, fused=True)
plan_u = cm.RcpsGraph(out, lab, cfg, fused=False)


def unfused():
    rcps.miss_counts(out, lab, lam_dev, counts=counts, totals=totals, zero=False)


def series(fn, k):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


for name, fn in (("hist kernel (unfused)", unfused), ("fused graph replay", lambda: plan_f.run()),
                 ("unfused graph replay (3 kernels + memsets)", lambda: plan_u.run())):
    torch.cuda.synchronize(); time.sleep(3.0)
    c0 = clk()
    cold = series(fn, 20)
    c1 = clk()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 2.0:
        series(fn, 50)
    c2, p2 = clk(), pw()
    hot = series(fn, 20)
    print(f"{name}: cold first5 {[round(x,3) for x in cold[:5]]} mean20 {sum(cold)/20:.4f} ms (clk {c0}->{c1} MHz) | "
          f"after 2 s of load: mean20 {sum(hot)/20:.4f} ms (clk {c2} MHz, {p2:.0f} W)")
