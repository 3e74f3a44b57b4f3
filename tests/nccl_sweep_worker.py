"""Worker for tests/test_rcps_gpu.py::test_two_rank_nccl_sweep (launched by torch.distributed.run, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import load_golden  # noqa: E402
from im2im_uq_b200.calibration import calibrate_model as cm  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    for case in ("fastmri_small", "temca_small", "top_risk_zero", "never_stops", "nasty_ragged"):
        g = load_golden(case)
        n = g["outputs"].shape[0]
        cuts = [n * r // world for r in range(world + 1)]
        lo, hi = cuts[rank], cuts[rank + 1]
        cfg = dict(g["config"], device=str(dev))
        out = torch.from_numpy(g["outputs"][lo:hi]).to(dev); lab = torch.from_numpy(g["labels"][lo:hi]).to(dev)
        lhat, stop, counts, visited = cm.rcps_sweep(out, lab, cfg, group=dist.group.WORLD)
        assert stop == int(g["stop_idx"]), (case, rank, stop)
        assert np.float32(lhat.numpy()) == g["lhat"]
        assert np.array_equal(counts.cpu().numpy(), g["counts_prime"][lo:hi])
    # the drop-in entry point with a sharded dataset: identity model, so dataset inputs are the head outputs
    from im2im_uq_b200.models.add_uncertainty import ModelWithUncertainty
    from im2im_uq_b200.models.quantile_layer import (quantile_regression_loss_fn,
                                                     quantile_regression_nested_sets_from_output)

    class _Id(torch.nn.Module):
        def forward(self, x):
            return x

    for case in ("fastmri_small", "batch65"):
        g = load_golden(case)
        n = g["outputs"].shape[0]
        cfg = dict(g["config"], device=str(dev))
        model = ModelWithUncertainty(_Id(), _Id(), quantile_regression_loss_fn,
                                     quantile_regression_nested_sets_from_output, cfg)
        ds = torch.utils.data.TensorDataset(torch.from_numpy(g["outputs"].copy()), torch.from_numpy(g["labels"].copy()))
        model, table = cm.calibrate_model(model, ds, cfg, group=dist.group.WORLD)
        lo, hi = n * rank // world, n * (rank + 1) // world
        assert np.float32(model.lhat.numpy()) == g["lhat"], (case, rank)
        assert np.array_equal(table.numpy(), g["calib_loss_table"][lo:hi]), (case, rank)
        model, full = cm.calibrate_model(model, ds, cfg, group=dist.group.WORLD, gather_table=True)
        assert np.array_equal(full.numpy(), g["calib_loss_table"]), (case, rank)     # the reference's whole table, on every rank
    # the captured plan: all-reduce fused with the decision over peer memory when symmetric memory works here, NCCL otherwise;
    # both must reproduce the reference's lhat / table rows, replay after replay
    for p2p, fused in ((True, True), (True, False), (False, False)):
        for case in ("fastmri_small", "temca_small", "top_risk_zero", "never_stops"):
            g = load_golden(case)
            n = g["outputs"].shape[0]
            cuts = [n * r // world for r in range(world + 1)]
            lo, hi = cuts[rank], cuts[rank + 1]
            cfg = dict(g["config"], device=str(dev))
            out = torch.from_numpy(g["outputs"][lo:hi]).to(dev); lab = torch.from_numpy(g["labels"][lo:hi]).to(dev)
            plan = cm.RcpsGraph(out, lab, cfg, group=dist.group.WORLD, n_total=n, p2p=p2p, fused=fused)
            if rank == 0 and case == "fastmri_small":
                print("PLAN_PATH", "fused single launch over peer memory" if plan.fused else
                      ("decide_p2p over peer memory" if plan.peer is not None else "nccl all-reduce"),
                      "kernels/replay", plan.kernels_per_replay, flush=True)
            for _ in range(3):
                lhat, stop, decided = plan.run()
                if not decided:
                    lhat, stop = plan.replay_on_host()
                assert stop == int(g["stop_idx"]), (case, p2p, rank, stop)
                assert np.float32(lhat.numpy()) == g["lhat"]
                assert np.array_equal(plan.table.cpu().numpy(), g["calib_loss_table"][lo:hi]), (case, p2p, rank)
                assert np.array_equal(plan.totals.cpu().numpy(), g["counts_prime"].sum(0, dtype=np.int64)), (case, p2p)
            plan.close()
    # images that straddle thread blocks, sharded unevenly over the ranks: fused peer path vs the single-GPU separate kernels
    from conftest import synth_scores
    n = 301
    out_all, lab_all = synth_scores(9, n, 1, 64, 64, device=dev)
    cfg = dict(uncertainty_type="quantiles", minimum_lambda=0.0, maximum_lambda=6.0, num_lambdas=1000, alpha=0.1,
               delta=0.1, device=str(dev), dataset="synthetic", rcps_loss="fraction_missed")
    ref = cm.RcpsGraph(out_all, lab_all, cfg, fused=False)
    want = ref.run()
    lo, hi = n * rank // world, n * (rank + 1) // world
    plan = cm.RcpsGraph(out_all[lo:hi].contiguous(), lab_all[lo:hi].contiguous(), cfg, group=dist.group.WORLD, n_total=n)
    for _ in range(4):
        got = plan.run()
        torch.cuda.synchronize()
        assert got[1] == want[1] and got[2] == want[2], (rank, got, want)
        assert torch.equal(plan.totals, ref.totals) and torch.equal(plan.counts, ref.counts[lo:hi])
        assert torch.equal(plan.table, ref.table[lo:hi])
    if rank == 0:
        print("FUSED_MULTI_GPU", "fused" if plan.fused else "not fused", flush=True)
    plan.close(); ref.close()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()          # graphs that captured NCCL kernels are gone (close()): teardown returns
    if rank == 0:
        print("NCCL_SWEEP_OK", flush=True)


if __name__ == "__main__":
    main()
