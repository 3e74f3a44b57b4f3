"""Standalone bandwidth of the BatchNorm/ReLU passes of the training step (one GPU): how far each is from the HBM roofline
when nothing else runs next to it.  usage: python tools/bn_bwd_bench.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200 import _lib  # noqa: E402


def timed(fn, reps=10):
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
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print(f"batch {B}; GB/s = algorithmic bytes / time (tensors larger than L2 except the deepest)")
    for side, C in ((320, 64), (160, 128), (80, 256), (40, 512)):
        n_pix = B * side * side
        z = torch.randn(B, side, side, C, device=dev).to(torch.bfloat16)
        dy = torch.randn(B, side, side, C, device=dev).to(torch.bfloat16)
        dz = torch.empty_like(z)
        y = torch.empty_like(z)
        gamma = torch.rand(C, device=dev) + 0.5
        beta = torch.randn(C, device=dev) * 0.1
        mean = torch.randn(C, device=dev) * 0.1
        rstd = torch.rand(C, device=dev) + 0.5
        sums = torch.zeros(2 * C, device=dev)
        T = n_pix * C * 2 / 1e9   # GB per pass over one tensor

        def bwd():
            _lib.check(lib.im2im_bn_relu_bwd_bf16(dy.data_ptr(), z.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                  mean.data_ptr(), rstd.data_ptr(), n_pix, C, sums.data_ptr(),
                                                  dz.data_ptr(), st), "bwd")

        def apply():
            _lib.check(lib.im2im_bn_apply_relu_bf16(z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), n_pix, C,
                                                    y.data_ptr(), st), "apply")

        def stats():
            _lib.check(lib.im2im_channel_stats_bf16(z.data_ptr(), n_pix, C, sums.data_ptr(), st), "stats")

        t_b, t_a, t_s = timed(bwd), timed(apply), timed(stats)
        print(f"  {side:3d}x{side:<3d} C={C:3d}  tensor {T*1e3:7.1f} MB | bn_relu_bwd (reduce 2 passes + apply 3 passes) "
              f"{t_b:6.3f} ms {5*T/t_b*1e3:6.0f} GB/s | bn_apply_relu (2 passes) {t_a:6.3f} ms {2*T/t_a*1e3:6.0f} GB/s | "
              f"channel_stats (1 pass) {t_s:6.3f} ms {T/t_s*1e3:6.0f} GB/s")
    del flush


if __name__ == "__main__":
    main()
