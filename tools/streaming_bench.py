"""Head convolution -> planes -> miss counts (two kernels, 12 B/pixel written and read back) versus the head convolution
whose epilogue books the ranks itself (im2im_head_conv3x3_tc_hist + im2im_rcps_counts_from_hist).  One GPU.
usage: python tools/streaming_bench.py [batch] [side] [n_lambdas]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200 import rcps  # noqa: E402
from im2im_uq_b200.conv import head_conv_tc, head_conv_tc_hist, pack_conv_weight, pad_head_weight  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 78
    side = int(sys.argv[2]) if len(sys.argv) > 2 else 320
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    m = torch.zeros(B, side, side, 64)
    m[..., :32] = torch.randn(B, side, side, 32, generator=g)
    m = m.to(dev).to(torch.bfloat16)
    hw = (torch.randn(3, 32, 3, 3, generator=g) * 0.08).to(dev)
    hb = torch.tensor([-0.4, 0.0, 0.4], device=dev)
    packed = pack_conv_weight(pad_head_weight(hw))
    planes = head_conv_tc(m, packed, hb, 3)
    labels = (planes[:, 1:2] + 0.6 * torch.randn(B, 1, side, side, generator=g).to(dev)).contiguous()
    lam = torch.linspace(0.0, 6.0, L, device=dev)
    counts = torch.zeros((B, L), dtype=torch.int32, device=dev)
    totals = torch.zeros(L, dtype=torch.int64, device=dev)
    hist = torch.zeros((B, L + 1), dtype=torch.int32, device=dev)
    want, _ = rcps.miss_counts(planes.view(B, 3, 1, side, side), labels, lam)

    t_head = timed(lambda: head_conv_tc(m, packed, hb, 3))
    t_count = timed(lambda: rcps.miss_counts(planes.view(B, 3, 1, side, side), labels, lam, counts=counts, totals=totals))
    t_hist = timed(lambda: head_conv_tc_hist(m, packed, hb, labels, lam, hist))
    hist.zero_()
    t_both = timed(lambda: (head_conv_tc_hist(m, packed, hb, labels, lam, hist), rcps.counts_from_hist(hist, counts, totals)))
    assert torch.equal(counts, want)
    px = B * side * side
    print(f"B={B} {side}x{side} L={L}")
    print(f"  head conv -> planes                    {t_head:8.3f} ms")
    print(f"  miss counts of the planes              {t_count:8.3f} ms   (two-stage total {t_head + t_count:.3f} ms, "
          f"{px * 12 * 2 / 1e6:.0f} MB of head tensor written + read)")
    print(f"  head conv with the histogram epilogue  {t_hist:8.3f} ms")
    print(f"  ... + counts from the histograms       {t_both:8.3f} ms   (streaming total; counts bit-identical)")


if __name__ == "__main__":
    main()
