#!/usr/bin/env python
"""Benchmark of the hot path: RCPS calibration images/sec on a 10k-image 320x320 calibration set (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  the reference's algorithm on the host CPU cores

A "step" is one complete calibration of the whole set with the scores already resident in HBM: zero the outputs, the
one-pass miss-count kernel over every lambda, (N>1: one NCCL all-reduce of the per-lambda totals), the host replay of
the reference's stopping rule (device->host copy of the totals + any replayed columns), and the fp32 loss-table
kernel.  The set is split evenly over the ranks (strong scaling, as BASELINE.json's north_star asks: "10k-image set,
1->8 GPUs").  `e2e` times the public entry point `calibrate_from_outputs` on HOST (pinned) score tensors, so the
host->device copy of the scores and the device->host copy of the loss table are inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "calibration images/sec (320x320, 10k set)"
UNIT = "images/s"

# BASELINE.json `configs`.  C3 is the configuration the metric is quoted on and the default.  The lambda
# grids are the reference's: fastmri [0,6]x1000 (experiments/fastmri_test/config.yml:29,37-39), temca [7,10]x100
# (experiments/temca_test/config.yml:29,37-39); alpha = delta = 0.1 in both (:24-27).  `noise` scales the synthetic label
# noise so that lambda-hat lands mid-grid on that grid.
CONFIGS = {
    "C2": dict(images=1000, side=320, lambdas=1000, lam_min=0.0, lam_max=6.0, noise=1.0, gpus=1, name="fastmri_test, 1k calibration images"),
    "C3": dict(images=10000, side=320, lambdas=1000, lam_min=0.0, lam_max=6.0, noise=1.0, gpus=8, name="fastmri_test"),
    "C4": dict(images=4000, side=512, lambdas=100, lam_min=7.0, lam_max=10.0, noise=4.6, gpus=4, name="temca_test"),
    "C5": dict(images=50000, side=640, lambdas=1000, lam_min=0.0, lam_max=6.0, noise=1.0, gpus=8, name="stress"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS), help="BASELINE.json configuration (default C3 = the "
                    "one the metric is quoted on: 10k x 320x320, lambda grid [0,6]x1000)")
    ap.add_argument("--images", type=int, default=None, help="calibration images in the whole job (default: the config's)")
    ap.add_argument("--side", type=int, default=None)
    ap.add_argument("--lambdas", type=int, default=None)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=192, help="images in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the calibration step kernel by kernel instead of "
                    "replaying its CUDA graph")
    ap.add_argument("--no-unet", action="store_true", help="skip the UNet forward / train-step sub-benchmark")
    ap.add_argument("--unet-batch", type=int, default=78, help="images per GPU per UNet step (78 = the reference's "
                    "batch size, experiments/fastmri_test/config.yml:45)")
    ap.add_argument("--unet-steps", type=int, default=6)
    ap.add_argument("--soak-seconds", type=float, default=0.0, help="run untimed steps for this long before the timed region "
                    "(power-capped steady state instead of the burst)")
    ap.add_argument("--no-unet-reference", action="store_true", help="skip the torch/cuDNN and CPU reference legs of the "
                    "UNet sub-benchmark")
    ap.add_argument("--e2e-images", type=int, default=None, help="images per GPU in the e2e (host buffers) leg; default: "
                    "the whole shard, capped at 16 GB of pinned host memory per rank")
    args = ap.parse_args()
    c = CONFIGS[args.config]
    args.images = args.images if args.images is not None else c["images"]
    args.side = args.side if args.side is not None else c["side"]
    args.lambdas = args.lambdas if args.lambdas is not None else c["lambdas"]
    args.lam_min, args.lam_max, args.noise = c["lam_min"], c["lam_max"], c["noise"]
    return args


def config_dict(args, device):
    return dict(alpha=0.1, delta=0.1, device=device, uncertainty_type="quantiles", minimum_lambda=args.lam_min,
                maximum_lambda=args.lam_max, num_lambdas=args.lambdas, rcps_loss="fraction_missed", dataset="synthetic",
                batch_size=78, q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)


def workload_name(args):
    return f"{args.config} {CONFIGS[args.config]['name']} RCPS calibration: {args.images} x 1x{args.side}x{args.side} fp32 " \
           f"quantile-head outputs, lambda grid [{args.lam_min:g},{args.lam_max:g}]x{args.lambdas}, alpha=delta=0.1"


def synth(n, side, device, seed, noise=1.0):
    """SURVEY.md §8c probe recipe (lambda-hat lands mid-grid; `noise` moves it for the temca grid); generated on `device`."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    shape = (n, 1, side, side)
    out = torch.empty((n, 3, 1, side, side), dtype=torch.float32, device=device)
    lab = torch.empty(shape, dtype=torch.float32, device=device)
    step = max(1, min(n, 512))
    for lo in range(0, n, step):  # chunked so temporaries stay small next to a 16 GB score tensor
        hi = min(n, lo + step)
        s = (hi - lo, 1, side, side)
        pred = torch.rand(s, generator=g, device=device)
        sig = 0.02 + 0.1 * torch.rand(s, generator=g, device=device)
        out[lo:hi, 0] = pred - sig * (0.5 + torch.rand(s, generator=g, device=device))
        out[lo:hi, 1] = pred
        out[lo:hi, 2] = pred + sig * (0.5 + torch.rand(s, generator=g, device=device))
        lab[lo:hi] = pred + noise * sig * torch.randn(s, generator=g, device=device)
    return out, lab


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (B200_PROFILING.md) - through NVML in this process
    (the library behind nvidia-smi) from a thread, every ~1 ms, so that even a 50 ms timed region is covered by dozens of
    samples; `nvidia-smi -lms` (one sample per ~30 ms) is the fallback when pynvml cannot be used."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index=0):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index
        self.thread = None
        self.rows = []
        self._stop = threading.Event()

    def _nvml_loop(self, nv, handle):
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(handle) / 1e3,
                                  int(reasons_fn(handle))))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.001)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: match by PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(self.gpu_index).pci_bus_id
            handle = None
            for i in range(nv.nvmlDeviceGetCount()):
                hnd = nv.nvmlDeviceGetHandleByIndex(i)
                if nv.nvmlDeviceGetPciInfo(hnd).bus == bus:
                    handle = hnd
            if handle is None:
                handle = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            rows = list(self.rows)
            if rows:
                sm = sorted(r[0] for r in rows)
                mask = 0
                for r in rows:
                    mask |= r[2]
                out = {"sm_mhz": float(sm[len(sm) // 2]), "sm_min_mhz": float(sm[0]), "sm_max_mhz": self.max_mhz,
                       "reasons": [nm for nm, bit in self.REASONS if mask & bit], "samples": len(rows),
                       "power_w_max": max(r[1] for r in rows), "source": "NVML in-process, ~1 ms period"}
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[1]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [nm for k, nm in enumerate(names) if any("Active" in r[4 + k] and "Not" not in r[4 + k] for r in rows)]
            out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][2]) if rows else None,
                   "reasons": reasons, "samples": len(rows), "source": "nvidia-smi -lms 20",
                   "power_w_max": max(float(r[3]) for r in rows) if rows else None}
        except Exception as e:  # never let monitoring break the benchmark
            out["error"] = repr(e)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy, of measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# --------------------------------------------------------------------------------------------- reference / CPU leg
def cpu_reference_sweep(sample_out, sample_lab, cfg, stop_full, n_full):
    """The reference's algorithm (calibrate_model.py:130-145): one full pass over the sample PER visited lambda, fp32
    mean, HB bound - restated in C (oracle port, all host threads).  The number of visited steps is pinned to what
    the full workload visits (columns L-1 .. stop_full, or its own early stop with the bound at the full set size) so
    the sample costs what its share of the real job costs."""
    import torch
    from oracle import rcps_oracle as orc
    lambdas = torch.linspace(cfg["minimum_lambda"], cfg["maximum_lambda"], cfg["num_lambdas"])
    dlambda = lambdas[1] - lambdas[0]
    px = float(sample_out[0, 0].size)
    t0 = time.perf_counter()
    steps = 0
    t_bound = 0.0
    first = 0 if stop_full is None else max(stop_full, 0)
    for j in reversed(range(first, cfg["num_lambdas"])):
        counts = orc.c_miss_counts(sample_out, sample_lab, float(lambdas[j] - dlambda))
        losses = torch.from_numpy(counts.astype("float32")) / px
        rhat = losses.mean()
        tb = time.perf_counter()
        rhat_plus = orc.hb_mu_plus(rhat.item(), n_full, cfg["delta"])
        t_bound += time.perf_counter() - tb
        steps += 1
        # stop_full=None: the reference's own early stop, with the bound evaluated at the FULL set size so the sample
        # visits the lambda steps the whole job would (risk per lambda is a population quantity)
        if stop_full is None and (rhat >= cfg["alpha"] or rhat_plus > cfg["alpha"]):
            break
    return time.perf_counter() - t0, steps, t_bound


def full_job_images_per_s(dt, t_bound, sample_images, n_full):
    """Throughput of the WHOLE job extrapolated from a sample: the data passes scale with the number of images, the
    Hoeffding-Bentkus solves (one per visited lambda step) do not - on a small sample they would otherwise dominate and
    understate the reference."""
    data = (dt - t_bound) * (n_full / float(sample_images))
    return n_full / (data + t_bound)


def run_reference_arm(args):
    """bench.py --impl reference: rank 0 only; K steps, each a bounded sample of the workload on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is the reference on ALL the host cores it can use
    n_threads = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(n_threads)      # read by libgomp when the oracle library is loaded below
    import numpy as np
    import torch
    torch.set_num_threads(n_threads)
    from oracle import rcps_oracle as orc
    orc.build()
    cfg = config_dict(args, "cpu")
    s_max = max(8, args.cpu_sample // 4)
    out, lab = synth(s_max, args.side, "cpu", 1234, args.noise)
    out, lab = out.numpy(), lab.numpy()
    stop_full = None  # early stop decided on the sample with the HB bound at the full set size
    for _ in range(max(args.warmup, 1)):
        cpu_reference_sweep(out[:2], lab[:2], cfg, args.lambdas - 3, args.images)
    # size the per-step sample so that the K timed steps end within ~100 s whatever K is: one probe sweep over 4 images
    # gives the cost per image (the sweep's cost is linear in images)
    probe_dt, _, probe_tb = cpu_reference_sweep(out[:4], lab[:4], cfg, stop_full, args.images)
    per_image = (probe_dt - probe_tb) / 4.0
    s = int(min(s_max, max(2, (100.0 / max(args.steps, 1) - probe_tb) / max(per_image, 1e-6))))
    out, lab = out[:s], lab[:s]
    times, bounds = [], []
    for _ in range(args.steps):
        dt, visited, tb = cpu_reference_sweep(out, lab, cfg, stop_full, args.images)
        times.append(dt)
        bounds.append(tb)
    total = sum(times)
    value = full_job_images_per_s(total, sum(bounds), s * args.steps, args.images * args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "inputs": "host memory"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                             "sample": f"{s} images x {visited} visited lambda steps per timed step (one full pass per "
                                       f"lambda step + fp32 mean + HB bound), C/OpenMP restatement of the reference loop; "
                                       f"value = whole-job throughput: data passes scaled to {args.images} images, the "
                                       f"{visited} HB solves counted once"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- UNet sub-benchmark
def _reference_quantile_loss(pred, target, params):
    """The reference's training loss as it is written there (quantile_layer.py:23-32 over pinball.py:12-24): boolean-mask
    assignment (each one a nonzero + host sync) - used ONLY by the torch reference legs below."""
    import torch

    def pinball(out, tgt, q):
        err = out - tgt
        loss = torch.zeros_like(tgt)
        neg, pos = err < 0, err > 0
        loss[neg] = q * err[neg].abs()
        loss[pos] = (1 - q) * err[pos].abs()
        return loss.mean()

    t = target.squeeze()
    return (params["q_lo_weight"] * pinball(pred[:, 0].squeeze(), t, params["q_lo"]) +
            params["q_hi_weight"] * pinball(pred[:, 2].squeeze(), t, params["q_hi"]) +
            params["mse_weight"] * torch.nn.functional.mse_loss(pred[:, 1].squeeze(), t))


def run_unet_reference_legs(args, dev, params, B, side, with_cpu):
    """What the reference's own code path costs on this box: its module graph (fp32 nn.Conv2d / BatchNorm2d / ReLU / MaxPool2d /
    bilinear Upsample, the mask-assignment pinball loss, torch.optim.Adam - core/scripts/train.py:147-165) through the
    libraries torch 2.11 dispatches to (cuDNN, TF32 convolutions by default), plus bf16 autocast + channels_last as the
    strongest library configuration, plus (N = 1 only) the same step on the host cores.  None of our kernels run here."""
    import torch
    from im2im_uq_b200.models.add_uncertainty import add_uncertainty
    from im2im_uq_b200.models.unet import UNet

    def build(device):
        torch.manual_seed(0)
        m = add_uncertainty(UNet(1, 1), params).to(device)
        m.use_native_inference = False        # the torch module graph, not our engines
        m.use_native_training = False
        return m

    def timed(fn, iters, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    g = torch.Generator(device=dev).manual_seed(5)
    out = {"what": "the reference's module graph through torch %s + cuDNN on the same GPU (library kernels only), batch %d, "
                   "1x%dx%d" % (torch.__version__, B, side, side)}
    import warnings
    for name, autocast, channels_last in (("torch_cudnn_tf32", False, False), ("torch_bf16_channels_last", True, True)):
        leg = {"precision": "bf16 autocast, channels_last" if autocast else
               "fp32 modules, cuDNN TF32 convolutions (torch default: cudnn.allow_tf32=%s)" % torch.backends.cudnn.allow_tf32}
        b = B
        while b >= 1:
            try:
                x = torch.randn(b, 1, side, side, device=dev, generator=g)
                y = x + 0.3 * torch.randn(b, 1, side, side, device=dev, generator=g)
                model = build(dev)
                if channels_last:
                    model = model.to(memory_format=torch.channels_last)
                    x = x.contiguous(memory_format=torch.channels_last)

                def fwd():
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                        return model(x)

                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    model.eval()
                    f_ms = timed(fwd, 3)
                    model.train()
                    opt = torch.optim.Adam(model.parameters(), lr=1e-4)

                    def train():
                        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                            pred = model(x)
                        loss = _reference_quantile_loss(pred.float(), y, params)
                        loss.item()
                        opt.zero_grad()
                        loss.backward()
                        opt.step()

                    t_ms = timed(train, 3)
                leg.update(batch=b, forward_ms=f_ms, forward_images_per_s=b / (f_ms * 1e-3), train_ms_per_step=t_ms,
                           train_images_per_s=b / (t_ms * 1e-3))
                del model, opt, x, y
                torch.cuda.empty_cache()
                break
            except torch.cuda.OutOfMemoryError:
                leg.setdefault("oom_at_batch", []).append(b)
                model = opt = x = y = None
                torch.cuda.empty_cache()
                b //= 2
        out[name] = leg
    if with_cpu:
        n_threads = len(os.sched_getaffinity(0))
        torch.set_num_threads(n_threads)
        bc = 4
        gc = torch.Generator().manual_seed(5)
        x = torch.randn(bc, 1, side, side, generator=gc)
        y = x + 0.3 * torch.randn(bc, 1, side, side, generator=gc)
        model = build("cpu")
        model.eval()
        with torch.no_grad():
            model(x)
            t0 = time.perf_counter()
            model(x)
            f_s = time.perf_counter() - t0
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        t0 = time.perf_counter()
        loss = _reference_quantile_loss(model(x), y, params)
        loss.item()
        opt.zero_grad()
        loss.backward()
        opt.step()
        t_s = time.perf_counter() - t0
        out["cpu"] = {"cores": n_threads, "batch": bc, "forward_images_per_s": bc / f_s, "train_images_per_s": bc / t_s,
                      "sample": "one forward and one training step at batch %d on the host cores (torch CPU, fp32)" % bc}
    return out


def run_unet_bench(args, world, rank, dev, group):
    """Second half of BASELINE.json's metric: UNet(1,1)+quantile-head images/s at 320x320, forward (inference engine)
    and full train step (forward, fused loss, backward, NCCL all-reduce of the flat fp32 gradients, fused Adam) on the
    native sm_100a kernels.  Per-GPU batch is fixed (weak scaling); images/s is the whole-job aggregate."""
    import torch
    import torch.distributed as dist
    from im2im_uq_b200.models.add_uncertainty import add_uncertainty
    from im2im_uq_b200.models.unet import UNet
    from im2im_uq_b200.models.unet_train import FusedAdam
    params = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)
    torch.manual_seed(0)
    model = add_uncertainty(UNet(1, 1), params).to(dev)
    B, side = args.unet_batch, args.side
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn(B, 1, side, side, device=dev, generator=g)
    y = x + 0.3 * torch.randn(B, 1, side, side, device=dev, generator=g)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    model.eval()
    with torch.no_grad():
        fwd_ms = timed(lambda: model(x), args.unet_steps)
        model.native_precision = "tf32"          # the reference's precision: fp32 activations, tcgen05 kind::tf32
        fwd_tf32_ms = timed(lambda: model(x), args.unet_steps)
        model.native_precision = "bf16"
    model.train()
    opt = FusedAdam(model.parameters(), lr=1e-4)
    last = {}

    def train_step():
        opt.zero_grad()
        loss = model.loss_fn(model(x), y)
        loss.backward()
        if world > 1:
            opt.gather_grads()
            dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM, group=group)  # 69 MB fp32 over NCCL
        opt.step(grad_scale=1.0 / world)
        last["loss"] = loss.item()  # the reference reads the loss every step (train.py:155)

    train_eager_ms = timed(train_step, args.unet_steps)

    # the same iteration captured into ONE CUDA graph (GraphedTrainStep); every step copies a fresh batch from pinned
    # host memory (H2D inside the timed region) and reads the loss back, like the reference's loop does
    from im2im_uq_b200.models.unet_train import GraphedTrainStep
    graphed = GraphedTrainStep(model, opt, x, y, group=group if world > 1 else None)
    x_host, y_host = x.cpu().pin_memory(), y.cpu().pin_memory()

    def train_step_graph():
        last["loss"] = graphed(x_host, y_host).item()

    train_ms = timed(train_step_graph, args.unet_steps)
    kernels_per_step = graphed.kernels_per_replay
    graphed.close()      # drop the captured graph (it holds NCCL kernels when data parallel) before anything else runs

    # calibrate_model end to end (BASELINE configs[1]: 1k calibration images, full UNet): host dataset -> native UNet
    # inference in batches -> scores stay in HBM -> one-pass RCPS -> lhat + loss table back on the host
    cal = None
    if rank == 0:
        from im2im_uq_b200.calibration.calibrate_model import calibrate_model
        n_cal = 1000
        gx = torch.Generator().manual_seed(7)
        xs = torch.randn(n_cal, 1, side, side, generator=gx).pin_memory()
        ys = (xs + 0.3 * torch.randn(n_cal, 1, side, side, generator=gx)).pin_memory()
        ds = torch.utils.data.TensorDataset(xs, ys)
        cfg = config_dict(args, str(dev))
        cfg.update(minimum_lambda=0.0, maximum_lambda=60.0, batch_size=50)
        model.eval()
        import contextlib, io
        def timed_calibration(c):
            calibrate_model(model, ds, c)            # warm-up (engine build, allocator)
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats(dev)
            base = torch.cuda.memory_allocated(dev)
            runs = []
            for _ in range(3):                       # wall clock around a 0.2 s host-driven call: report the median
                t0 = time.perf_counter()
                _, tab = calibrate_model(model, ds, c)
                torch.cuda.synchronize()
                runs.append(time.perf_counter() - t0)
            return sorted(runs)[1], runs, (torch.cuda.max_memory_allocated(dev) - base) / 2**20, float(model.lhat), tab

        with contextlib.redirect_stdout(io.StringIO()):
            dt, runs, peak_mib, lhat_s, tab_s = timed_calibration(cfg)                    # streaming: no (N,3,C,H,W) tensor
            dt2, runs2, peak2_mib, lhat_2, tab_2 = timed_calibration(dict(cfg, streaming_calibration=False))
        assert lhat_s == lhat_2 and torch.equal(tab_s, tab_2), "streaming and two-stage calibration disagree"
        cal = {"images": n_cal, "seconds": dt, "images_per_s": n_cal / dt, "runs_s": runs, "lhat": lhat_s,
               "peak_extra_hbm_mib": peak_mib,
               "two_stage": {"seconds": dt2, "images_per_s": n_cal / dt2, "runs_s": runs2, "peak_extra_hbm_mib": peak2_mib,
                             "note": "config['streaming_calibration'] = False: head outputs of the whole set kept in HBM, "
                                     "then one RCPS pass; same lhat and table bit for bit (asserted)"},
               "api": "core.calibration.calibrate_model.calibrate_model(model, dataset, config)",
               "note": "host TensorDataset -> H2D -> native UNet forward whose head epilogue books the RCPS ranks (no head "
                       "tensor) -> sweep on the (N, L) counts -> table D2H; random-init weights"}
        model.train()
    reference = None
    if rank == 0 and not args.no_unet_reference:
        del graphed, opt
        torch.cuda.empty_cache()
        reference = run_unet_reference_legs(args, dev, params, B, side, with_cpu=(world == 1))
    fwd_flop, train_flop = 125.29e9, 375.87e9  # conv 2*MACs per 320x320 image; train = 3x forward
    scale = (side / 320.0) ** 2
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peaks["bf16_tflops"])
        peak_sus = float(peaks.get("bf16_tflops_sustained", peak))
        src = "MEASURED_PEAKS.json bf16_tflops (burst) / bf16_tflops_sustained, of measured"
    except Exception:
        peak, peak_sus, src = 1590.0, 1400.0, "fallback 1.59 PFLOP/s burst / 1.4 sustained (B200_PROFILING.md)"
    fwd_tf = B * fwd_flop * scale / (fwd_ms * 1e-3) / 1e12
    train_tf = B * train_flop * scale / (train_ms * 1e-3) / 1e12
    return {"model": "UNet(1,1)+quantile head, 17.27M params", "image": f"1x{side}x{side}", "batch_per_gpu": B,
            "scaling": "weak", "dtype": "bf16 operands, fp32 accumulate/params",
            "precision": {"headline": "bf16 (forward_*, train_*): bf16 operands/activations, tcgen05 kind::f16, fp32 accumulation, "
                          "fp32 master weights / statistics / optimizer", "reference_precision_mode": "tf32 (forward_tf32_*): fp32 "
                          "activations, tcgen05 kind::tf32 - what torch does with the reference's fp32 modules on a GPU; "
                          "inference only"},
            "forward_images_per_s": world * B / (fwd_ms * 1e-3), "forward_ms": fwd_ms,
            "forward_tf32_images_per_s": world * B / (fwd_tf32_ms * 1e-3), "forward_tf32_ms": fwd_tf32_ms,
            "train_images_per_s": world * B / (train_ms * 1e-3), "train_ms_per_step": train_ms,
            "train_step": "one CUDA graph per step: H2D of the batch from pinned memory + forward + fused pinball/MSE "
                          "loss + backward + " + ("NCCL all-reduce of 69 MB fp32 grads + " if world > 1 else "") +
                          "fused Adam, then loss.item()",
            "train_ms_per_step_eager": train_eager_ms, "train_kernels_per_step": kernels_per_step,
            "final_loss": last.get("loss"), "calibrate_model_e2e": cal, "reference": reference,
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "peak": peak, "peak_sustained": peak_sus, "peak_source": src,
                         "forward_achieved": fwd_tf, "forward_frac": fwd_tf / peak,
                         "train_achieved": train_tf, "train_frac": train_tf / peak,
                         "train_frac_of_sustained": train_tf / peak_sus, "forward_frac_of_sustained": fwd_tf / peak_sus,
                         "forward_tf32_achieved": B * fwd_flop * scale / (fwd_tf32_ms * 1e-3) / 1e12,
                         "forward_tf32_frac_of_half_rate_peak": B * fwd_flop * scale / (fwd_tf32_ms * 1e-3) / 1e12 / (peak / 2),
                         "flops": "conv 2*MACs only: 125.29 GFLOP fwd, 375.87 GFLOP train per 320x320 image"}}


def probe_numa_node(torch, index: int):
    """sysfs does not say which NUMA node the GPU hangs off (virtualised PCI topology): measure it - for every node, run on
    its CPUs, pin a 256 MB buffer there (first touch at pin time) and time host->device copies; the fastest node wins when
    it is at least 10 % faster than the slowest.  ~0.3 s per node, outside every timed region."""
    try:
        import glob
        nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
        if len(nodes) < 2:
            return None
        all_cpus = os.sched_getaffinity(0)
        dev = torch.device("cuda", index)
        dst = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        rates = {}
        for nd in nodes:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{nd}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = cpus & all_cpus
            if not allowed:
                continue
            os.sched_setaffinity(0, allowed)
            src = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
            src.fill_(1)
            dst.copy_(src, non_blocking=True); torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize(dev)
            rates[nd] = 3 * (256 << 20) / (time.perf_counter() - t0)
            del src
        os.sched_setaffinity(0, all_cpus)
        if len(rates) < 2:
            return None
        best = max(rates, key=rates.get)
        return best if rates[best] > 1.1 * min(rates.values()) else None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_numa_node(torch, index: int):
    """Run this process on the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers (first touch) are
    allocated next to the GPU's PCIe root: the host->device copies of the e2e leg then do not cross the socket link.
    Best effort: returns a description of what was done, never raises."""
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev_id = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            node = probe_numa_node(torch, index)
            if node is None:
                return "numa node unknown (-1), single node or probe inconclusive: not bound"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return f"numa node {node}: none of its cpus is allowed here: not bound"
        os.sched_setaffinity(0, allowed)
        return f"bound to numa node {node} ({len(allowed)} cpus)"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


# --------------------------------------------------------------------------------------------- our arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from im2im_uq_b200 import _lib, rcps
    from im2im_uq_b200.calibration import calibrate_model as cm
    from im2im_uq_b200.calibration import sweep
    from im2im_uq_b200.models.add_uncertainty import ModelWithUncertainty
    from im2im_uq_b200.models.quantile_layer import (quantile_regression_loss_fn,
                                                     quantile_regression_nested_sets_from_output)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # Native libraries write to fd 1 behind Python's back (NCCL prints "NCCL version ..." at communicator init): point fd 1
    # at stderr for the whole run and keep the real stdout for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    _lib.load()

    cfg = config_dict(args, str(dev))
    cuts = [args.images * r // world for r in range(world + 1)]
    n_local = cuts[rank + 1] - cuts[rank]
    px = args.side * args.side
    out, lab = synth(n_local, args.side, dev, 1000 + rank, args.noise)
    lambdas, dlambda, lam_prime, default_lhat = sweep.lambda_grid(cfg)
    lam_dev = lam_prime.to(dev)
    L = args.lambdas
    counts = torch.empty((n_local, L), dtype=torch.int32, device=dev)
    totals = torch.empty((L,), dtype=torch.int64, device=dev)
    table = torch.empty((n_local, L), dtype=torch.float32, device=dev)
    k_start = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k_end = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def column_to_losses(col):
        return rcps.loss_table(col.reshape(-1, 1).contiguous(), px)[:, 0]

    result = {}
    decide = cm.device_decide_fn(px, cfg)

    def step(i=None):
        counts.zero_(); totals.zero_()  # torch memset kernels on the current stream (plumbing)
        if i is not None:
            k_start[i].record()
        rcps.miss_counts(out, lab, lam_dev, counts=counts, totals=totals, zero=False)
        if i is not None:
            k_end[i].record()
        if group is not None:
            dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)  # the one collective: int64[L] over NCCL
        res = decide(totals, args.images, read=False)                    # stop rule screened on the device
        rcps.loss_table(counts, px, out=table, first_visited_dev=res[3:])
        host = res.cpu()                                                 # 16-byte device->host read = the step's result
        stop, replayed = int(host[0]), 0
        if not bool(host[1]):  # a column fell inside the guard band: replay the reference's expression on the host
            stats = {}
            lhat_t, stop, visited = sweep.sweep_from_counts(counts, totals, px, cfg, column_to_losses, ascending=True,
                                                            group=group, stats=stats, n_total=args.images,
                                                            totals_already_reduced=True)
            rcps.loss_table(counts, px, first_visited_col=max(stop, 0), out=table)
            replayed = stats.get("replayed_columns")
        lhat = float(lambdas[stop]) if stop >= 0 else float(default_lhat)
        result.update(lhat=lhat, stop=stop, replayed=replayed)

    plan = None
    if not args.no_graph:
        # steady-state path: the step's device work captured once into a CUDA graph (cm.RcpsGraph) and replayed
        plan = cm.RcpsGraph(out, lab, cfg, group=group, n_total=args.images)

        def step(i=None):  # noqa: F811 - replaces the kernel-by-kernel step above
            lhat_t, stop, decided = plan.run(after_replay=k_end[i].record if i is not None else None)
            replayed = 0
            if not decided:
                stats = {}
                lhat_t, stop = plan.replay_on_host(stats)
                replayed = stats.get("replayed_columns")
            result.update(lhat=float(lhat_t), stop=stop, replayed=replayed)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    if args.soak_seconds > 0:
        # optional: keep stepping (untimed) for this long first, to measure the power-capped steady state instead of the
        # burst a single calibration really is
        sync_all()
        t_probe = time.perf_counter()
        for _ in range(5):
            step()
        sync_all()
        per_step = max((time.perf_counter() - t_probe) / 5, 1e-6)
        extra = torch.tensor([int(min(100000, args.soak_seconds / per_step))], dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(extra, src=0)     # the step is collective: every rank runs the same number of them
        for _ in range(int(extra)):
            step()
    sync_all()
    sampler = ClockSampler(local_rank)       # NVML thread, ~1 ms period: covers the timed region with dozens of samples
    if rank == 0:
        sampler.start()
    sync_all()
    launches0 = _lib.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_start.record()
    for i in range(args.steps):
        if plan is not None:
            k_start[i].record()     # graph path: the events bracket the replay = the step's kernel(s), inside the timed region
        step(i)
    t_end.record()
    sync_all()
    launches = _lib.launch_count() - launches0
    if plan is not None:
        launches += args.steps * plan.kernels_per_replay  # kernels replayed from the captured graph
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "the timed region (%d steps, %.1f ms)" % (args.steps, t_start.elapsed_time(t_end))
    ms_total = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    tail_stamps = plan.tail_stamps_us() if plan is not None else None
    if tail_stamps is not None and world > 1:
        # every rank's view of the last step's tail (when its last block finished streaming, when the peers had arrived ...)
        gathered = [None] * world
        dist.all_gather_object(gathered, tail_stamps)
        tail_stamps = {"per_rank": gathered}
    plan_fused = plan is not None and plan.fused
    plan_peer = plan is not None and plan.peer is not None
    plan_launches = plan.kernels_per_replay if plan is not None else None
    kernel_note = "CUDA events around rcps_hist_kernel in the kernel-by-kernel step, inside the timed region"
    if plan is not None:
        kernel_note = ("CUDA events around each graph replay inside the timed region; the replay is ONE kernel "
                       "(rcps_hist_kernel<fused>: counts + totals exchange + decision + loss table)") if plan_fused else \
                      "CUDA events around each graph replay inside the timed region (memsets + kernels of the multi-launch step)"
    kernel_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(k_start, k_end)) / args.steps],
                             dtype=torch.float64, device=dev)
    launches_t = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches_t, op=dist.ReduceOp.SUM)
    ms_per_step = float(ms_total) / args.steps
    value = args.images / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ e2e: public API on HOST buffers
    e2e = None
    if not args.no_e2e:
        class _Id(torch.nn.Module):
            def forward(self, x):
                return x
        model = ModelWithUncertainty(_Id(), _Id(), quantile_regression_loss_fn,
                                     quantile_regression_nested_sets_from_output, cfg)
        # host buffers: the whole shard when it fits 16 GB of pinned memory per rank, else a leading block of it
        per_image = px * 16
        n_e2e = args.e2e_images if args.e2e_images is not None else min(n_local, max(1, (16 << 30) // per_image))
        n_e2e = min(n_e2e, n_local)
        cuts_e = torch.tensor([n_e2e], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(cuts_e, op=dist.ReduceOp.SUM)
        n_e2e_total = int(cuts_e)
        all_cpus = os.sched_getaffinity(0)
        numa = bind_to_gpu_numa_node(torch, local_rank)      # pinned pages land on the GPU's NUMA node (first touch)
        host_out = torch.empty((n_e2e,) + tuple(out.shape[1:]), dtype=torch.float32, pin_memory=True)
        host_lab = torch.empty((n_e2e,) + tuple(lab.shape[1:]), dtype=torch.float32, pin_memory=True)
        host_out.copy_(out[:n_e2e]); host_lab.copy_(lab[:n_e2e])
        os.sched_setaffinity(0, all_cpus)                    # the CPU baseline below uses every core again
        torch.cuda.synchronize()
        if plan is not None:
            plan.close()                                     # releases its references to the score tensors
        del out, lab
        torch.cuda.empty_cache()

        def e2e_step():
            m, tbl = cm.calibrate_from_outputs(model, host_out, host_lab, cfg, group=group)
            return float(m.lhat), tbl  # tbl is a CPU tensor: the device->host read of the step's result

        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            lh, tbl = e2e_step()
        sync_all()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if n_e2e_total == args.images:
            assert abs(lh - result["lhat"]) == 0.0, (lh, result)
        e2e = {"value": n_e2e_total * args.e2e_steps / float(dt), "unit": UNIT,
               "h2d_bytes_per_step": n_e2e_total * px * 16 + L * 4 * world,
               "d2h_bytes_per_step": n_e2e_total * L * 4 + L * 8 * world, "steps": args.e2e_steps,
               "images": n_e2e_total,
               "sample": None if n_e2e_total == args.images else
               f"{n_e2e_total} of {args.images} images (pinned host memory capped at 16 GB per rank)",
               "host_buffers": "pinned, " + numa,
               "api": "im2im_uq_b200.calibration.calibrate_model.calibrate_from_outputs(model, outputs_cpu, labels_cpu, config)"}

    unet = None
    if not args.no_unet:
        if plan is not None:
            plan.close()
        try:
            del host_out, host_lab
        except NameError:
            pass
        try:
            del out, lab                     # --no-e2e: the score tensors are still alive
        except NameError:
            pass
        del counts, table, totals
        torch.cuda.empty_cache()
        unet = run_unet_bench(args, world, rank, dev, group)

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = n_local * px * 16 + n_local * L * 4
        achieved = alg_bytes / (float(kernel_ms) * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch are an ncu counter: read from the committed capture of this very shape, if there is one
            name = "r2_rcps_fused_ncu_summary.json" if plan_fused else "rcps_hist_ncu_summary.json"
            prof = json.load(open(os.path.join(ROOT, "profiles", name)))
            if prof.get("images") == n_local and prof.get("side") == args.side:
                traffic = prof.get("dram_bytes_per_launch")
                traffic_src = f"from profile: profiles/{name} (ncu --set full, dram__bytes_read.sum + " \
                              "dram__bytes_write.sum of one launch at this shape; not measured in this run)"
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args), "images_per_gpu": n_local,
                           "l2": "inputs (%.2f GB per GPU) exceed the 126 MB L2; no flush needed" % (n_local * px * 16 / 1e9),
                           "cuda_graph": plan is not None,
                           "launches_per_step": plan_launches, "fused_tail_us": tail_stamps,
                           "totals_allreduce": (("inside the one fused launch, over NVLink peer memory (im2im_rcps_calibrate_fused)"
                                                 if plan_fused else
                                                 "fused with the decision over NVLink peer memory (im2im_rcps_decide_p2p)")
                                                if plan_peer else ("NCCL" if world > 1 else "none (single GPU)")),
                           "lhat": result["lhat"], "lhat_index": result["stop"],
                           "replayed_columns": result["replayed"], "parallelism": (f"image shards x{world}, one exchange of the "
                           "uint64[L] totals per calibration (" + ("NVLink peer memory" if plan_peer else "NCCL all-reduce") + ")")
                           if world > 1 else "single GPU"},
                "roofline": {"bound": "hbm", "kernel": "rcps_hist_kernel<fused>" if plan_fused else "rcps_hist_kernel<staged>", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": float(kernel_ms),
                             "kernel_ms_how": kernel_note, "peak_source": peak_src},
                "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches_t), "unet": unet}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import rcps_oracle as orc
            orc.build()
            s_out, s_lab = synth(args.cpu_sample, args.side, "cpu", 1234, args.noise)
            dt, visited, tb = cpu_reference_sweep(s_out.numpy(), s_lab.numpy(), config_dict(args, "cpu"),
                                                  result["stop"], args.images)
            line["cpu_baseline"] = {"value": full_job_images_per_s(dt, tb, args.cpu_sample, args.images), "unit": UNIT,
                                    "cores": orc.num_threads(), "kind": "port",
                                    "sample": f"{args.cpu_sample} images x {visited} visited lambda steps (the steps the "
                                              f"full set visits; one full pass + fp32 mean + HB bound per step), "
                                              f"{dt:.1f} s of which {tb:.1f} s in the {visited} HB solves; value = whole-"
                                              f"job throughput (data passes scaled to {args.images} images, HB solves "
                                              f"counted once)"}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if plan is not None:
        plan.close()           # graphs that captured NCCL kernels / peer mappings are destroyed before the communicator (idempotent)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
