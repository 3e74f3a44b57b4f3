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
    # the captured plan: all-reduce fused with the decision over peer memory when symmetric memory works here, NCCL otherwise;
    # both must reproduce the reference's lhat / table rows, replay after replay
    for p2p in (True, False):
        for case in ("fastmri_small", "temca_small", "top_risk_zero", "never_stops"):
            g = load_golden(case)
            n = g["outputs"].shape[0]
            cuts = [n * r // world for r in range(world + 1)]
            lo, hi = cuts[rank], cuts[rank + 1]
            cfg = dict(g["config"], device=str(dev))
            out = torch.from_numpy(g["outputs"][lo:hi]).to(dev); lab = torch.from_numpy(g["labels"][lo:hi]).to(dev)
            plan = cm.RcpsGraph(out, lab, cfg, group=dist.group.WORLD, n_total=n, p2p=p2p)
            if p2p and rank == 0:
                print("P2P_PATH", "peer-memory" if plan.peer is not None else "nccl-fallback: " + getattr(plan, "p2p_error", "?"))
            for _ in range(3):
                lhat, stop, decided = plan.run()
                if not decided:
                    lhat, stop = plan.replay_on_host()
                assert stop == int(g["stop_idx"]), (case, p2p, rank, stop)
                assert np.float32(lhat.numpy()) == g["lhat"]
                assert np.array_equal(plan.table.cpu().numpy(), g["calib_loss_table"][lo:hi]), (case, p2p, rank)
                assert np.array_equal(plan.totals.cpu().numpy(), g["counts_prime"].sum(0, dtype=np.int64)), (case, p2p)
    del plan
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("NCCL_SWEEP_OK", flush=True)
    sys.stdout.flush()
    os._exit(0)   # skip NCCL teardown: ncclCommDestroy can wait on captured graphs (see bench.py)


if __name__ == "__main__":
    main()
