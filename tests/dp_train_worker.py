"""Worker for tests/test_unet_train_gpu.py::test_two_rank_data_parallel_step (torch.distributed.run, one rank per GPU).

Row b5 of SURVEY.md §8: the reference trains under nn.DataParallel (core/scripts/train.py:22-27,112-115): the batch is
scattered over the replicas, every replica normalises with ITS OWN BatchNorm batch statistics, the loss is the mean over the
gathered batch and the replicas' gradients are summed - i.e. the gradient is the average of the per-replica gradients of
the per-replica mean losses when the micro-batches are equal.  Here: one process per GPU, native engine on the local
micro-batch, ONE NCCL all-reduce of the flat fp32 gradient buffer, scaled by 1/world in the fused Adam.  Every rank also
evaluates the DataParallel arithmetic with the torch module graph (fp32) on the whole batch and compares."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200.models.add_uncertainty import add_uncertainty  # noqa: E402
from im2im_uq_b200.models.unet import UNet  # noqa: E402
from im2im_uq_b200.models.unet_train import FusedAdam  # noqa: E402

PARAMS = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=0.0, q_hi_weight=0.0, mse_weight=1.0)


def build(dev):
    torch.manual_seed(0)
    return add_uncertainty(UNet(1, 1), PARAMS).to(dev).train()


def dataparallel_reference_grads(x, y, world, dev, autocast=False):
    """What nn.DataParallel computes, written out: per-replica forward (own BN statistics), loss on the gathered batch.
    autocast=True evaluates the same arithmetic with torch's bf16 autocast: the yardstick for what bf16 operands cost."""
    m = build(dev)
    m.use_native_training = False
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            outs = [m(xs).float() for xs in x.chunk(world)]        # scatter + per-replica forward
        loss = m.loss_fn(torch.cat(outs, dim=0).cpu(), y.cpu())    # gather, loss on the full batch (torch formulas on CPU tensors)
        loss.backward()                                            # gradients summed over replicas
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}, float(loss)


def check_against_dataparallel(grads, ref, ref_ac):
    """Averaged native gradients vs the DataParallel arithmetic in fp32 (``ref``), with the same arithmetic under torch's
    bf16 autocast (``ref_ac``) as the yardstick: per layer, relative L2 error <= max(6e-2, 1.5 x autocast's error on that
    layer) and norms within 20 %.  A wrong data-parallel step - a missing 1/world (every norm off by 2x), the wrong loss
    normalisation, gradients of one replica only - fails both.  Returns the worst layer."""
    worst = ("", 0.0)
    for n, gr in ref.items():
        if gr.norm() < 1e-7 or n.endswith("double_conv.0.bias") or n.endswith("double_conv.3.bias"):
            continue                                               # conv biases in front of a BatchNorm: exactly zero here
        rel = float((grads[n] - gr).norm() / gr.norm())
        rel_ac = float((ref_ac[n] - gr).norm() / gr.norm())
        ratio = float(grads[n].norm() / gr.norm())
        assert rel <= max(6e-2, 1.5 * rel_ac), (n, rel, rel_ac)
        assert 0.8 <= ratio <= 1.25, (n, ratio)
        if rel > worst[1]:
            worst = (n, rel)
    return worst


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(11)
    B = 8 * world
    x = torch.randn(B, 1, 96, 96, generator=g).to(dev)
    y = (x.cpu() + 0.3 * torch.randn(B, 1, 96, 96, generator=g)).to(dev)
    ref, ref_loss = dataparallel_reference_grads(x, y, world, dev)
    ref_ac, _ = dataparallel_reference_grads(x, y, world, dev, autocast=True)

    model = build(dev)
    opt = FusedAdam(model.parameters(), lr=1e-3)
    xs, ys = x.chunk(world)[rank], y.chunk(world)[rank]
    opt.zero_grad()
    loss = model.loss_fn(model(xs), ys)
    assert "_native_train_engine" in model.__dict__
    loss.backward()
    opt.gather_grads()
    dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM)           # the one collective of the step
    grads, off = {}, 0
    for (n, p) in model.named_parameters():
        k = p.numel()
        grads[n] = (opt.flat_grad[off:off + k] / world).view_as(p).clone()
        off += k
    mean_loss = torch.tensor([float(loss)], device=dev)
    dist.all_reduce(mean_loss)
    assert abs(float(mean_loss) / world - ref_loss) <= 2e-3 * abs(ref_loss), (float(mean_loss) / world, ref_loss)
    worst = check_against_dataparallel(grads, ref, ref_ac)
    # the optimizer step uses the averaged gradient on every rank: parameters stay identical across ranks
    opt.step(grad_scale=1.0 / world)
    check = opt.flat_param.clone()
    dist.broadcast(check, src=0)
    assert torch.equal(check, opt.flat_param)
    if rank == 0:
        print(f"DP_TRAIN_OK world={world} loss {float(mean_loss) / world:.6f} (DataParallel arithmetic {ref_loss:.6f}); "
              f"worst per-layer gradient rel-L2 {worst[1]:.3e} ({worst[0]})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
