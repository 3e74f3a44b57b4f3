"""Dev check: the halo convolution as CTA pairs (cta_group::2, IM2IM_HALO_PAIR=1) against the single-CTA kernel.

The two modes accumulate every output element over the same K order, so the outputs (and the pooled tensor) must be
bit-identical; the fused statistics are fp32 atomics in a different order (compared with a tolerance).  The env switch is
read once per process, so each mode runs in its own subprocess:  python tools/halo_pair_check.py [bench_batch]
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # B, H, W, c1, c2, cout
    (1, 16, 8, 64, 0, 64),        # one tile: the pair's second tile lies behind the batch
    (3, 16, 8, 64, 0, 64),        # odd tile count
    (2, 32, 24, 64, 0, 64),
    (2, 32, 32, 64, 64, 64),      # concatenation (two tensor maps)
    (1, 48, 40, 64, 0, 128),      # N = 128: 64 weight rows per CTA
    (5, 64, 64, 128, 0, 64),
    (4, 160, 160, 64, 0, 64),     # more tiles than pairs: persistent loop, both TMEM buffers, ring wrap-around
    (2, 32, 32, 128, 0, 128),     # IM2IM_HALO_PAIR_WIDE: N = 128 blocks whose weights fit only when halved (two ring stages)
    (3, 48, 40, 128, 0, 256),
]


def worker(tag, bench_b):
    import torch
    from im2im_uq_b200.conv import conv_igemm, conv_igemm_stats, conv_igemm_pool
    dev = torch.device("cuda:0")
    out = {}
    for i, (B, H, W, c1, c2, co) in enumerate(CASES):
        g = torch.Generator(device="cpu").manual_seed(100 + i)
        x1 = torch.randn(B, H, W, c1, generator=g).to(dev).to(torch.bfloat16)
        x2 = torch.randn(B, H, W, c2, generator=g).to(dev).to(torch.bfloat16) if c2 else None
        w = (torch.randn(co, 9, c1 + c2, generator=g) / 24).to(dev).to(torch.bfloat16)
        b = torch.randn(co, generator=g).to(dev)
        out[f"plain{i}"] = conv_igemm(x1, w, b, True, x2).cpu()
        out[f"f32_{i}"] = conv_igemm(x1, w, b, False, x2, torch.float32).cpu()
        if co == 64:
            sums = torch.zeros(2 * co, device=dev)
            out[f"stats{i}"] = conv_igemm_stats(x1, w, 1, sums, x2=x2)[0].cpu()
            out[f"sums{i}"] = sums.cpu()
            z = torch.randn(B, H, W, co, generator=g).to(dev).to(torch.bfloat16)
            gamma = torch.rand(co, generator=g).to(dev) + 0.5; beta = torch.rand(co, generator=g).to(dev) - 0.5
            mean = torch.zeros(co, device=dev); rstd = torch.ones(co, device=dev)
            sums2 = torch.zeros(2 * co, device=dev)
            out[f"bwd{i}"] = conv_igemm_stats(x1, w, 2, sums2, x2=x2, bn=(z, gamma, beta, mean, rstd))[0].cpu()
            out[f"sums2_{i}"] = sums2.cpu()
        y, pooled = conv_igemm_pool(x1, w, b, True, x2)
        out[f"poolfull{i}"] = y.cpu()
        if pooled is not None:
            out[f"pool{i}"] = pooled.cpu()
    torch.cuda.synchronize()
    torch.save(out, f"/tmp/halo_pair_{tag}.pt")
    print(f"[{tag}] numerics cases done", flush=True)

    def t(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        e.record(); torch.cuda.synchronize()
        return a.elapsed_time(e) / iters

    B = bench_b
    for (H, c1, c2, co) in ((320, 64, 0, 64), (320, 64, 64, 64), (160, 64, 0, 128), (160, 128, 0, 64)):
        x1 = torch.randn(B, H, H, c1, device=dev).to(torch.bfloat16)
        x2 = torch.randn(B, H, H, c2, device=dev).to(torch.bfloat16) if c2 else None
        w = (torch.randn(co, 9, c1 + c2, device=dev) / 30).to(torch.bfloat16)
        sums = torch.zeros(2 * co, device=dev)
        fl = 2.0 * B * H * H * co * 9 * (c1 + c2)
        ms0 = t(lambda: conv_igemm(x1, w, x2=x2))
        line = f"[{tag}] B{B} {H}^2 {c1}+{c2}->{co}: plain {ms0:.3f} ms {fl / ms0 * 1e-9:.0f} TF"
        if co == 64:
            ms1 = t(lambda: conv_igemm_stats(x1, w, 1, sums, x2=x2))
            line += f" | stats fwd {ms1:.3f} ms {fl / ms1 * 1e-9:.0f} TF"
        print(line, flush=True)


def main():
    bench_b = sys.argv[1] if len(sys.argv) > 1 else "78"
    for tag, val in (("single", "0"), ("pair", "1")):
        env = dict(os.environ, IM2IM_HALO_PAIR=val, IM2IM_HALO_PAIR_WIDE=val)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", tag, bench_b], env=env, timeout=150)
        if r.returncode != 0:
            print(f"{tag}: worker failed rc={r.returncode}")
            sys.exit(1)
    import torch
    a, b = torch.load("/tmp/halo_pair_single.pt"), torch.load("/tmp/halo_pair_pair.pt")
    ok = True
    for k in sorted(a):
        if k.startswith("sums"):
            d = (a[k] - b[k]).abs().max().item(); s = a[k].abs().max().item()
            good = d <= 2e-4 * max(s, 1.0)
            print(f"{k}: max diff {d:.3g} (scale {s:.3g}) {'OK' if good else 'FAIL'}")
        else:
            good = torch.equal(a[k], b[k])
            if not good:
                d = (a[k].float() - b[k].float()).abs()
                print(f"{k}: NOT identical, max diff {d.max().item():.4g}, {int((d > 0).sum())} of {d.numel()} differ  FAIL")
        ok &= good
    print("PAIR == SINGLE:", "OK" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(sys.argv[2], int(sys.argv[3]))
    else:
        main()
