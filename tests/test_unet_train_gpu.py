"""GPU (-m gpu): the native training step (tcgen05 forward/dgrad/wgrad + fused BN/loss/Adam) end to end, driven through the
reference's own loop shape (forward -> loss_fn -> backward -> optimizer.step, core/scripts/train.py:152-162), against
PyTorch fp32 autograd through the same modules.

Tolerance rationale: operands/activations are bf16.  At random init the pinball loss has sign-valued gradients, so even
torch's own bf16 autocast deviates from its fp32 gradients by ~0.4 (median relative error) - that run is used as the
yardstick: the native engine must be at least as close to fp32 as autocast is, its loss must match to 1e-3, and the loss
trajectory over Adam steps must track torch's within 3 %."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PARAMS = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)


def _build():
    from core.models.add_uncertainty import add_uncertainty
    from core.models.trunks.unet import UNet
    torch.manual_seed(0)
    return add_uncertainty(UNet(1, 1), PARAMS).to("cuda:0").train()


def _grads(model, x, y, native, autocast=False):
    model.use_native_training = native
    model.zero_grad(set_to_none=True)
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = model(x)
        pred = pred.float()
    else:
        pred = model(x)
    loss = model.loss_fn(pred, y)
    loss.backward()
    # conv biases that feed a BatchNorm get no gradient from the native engine (exactly zero): None -> zeros
    return pred.detach(), loss.item(), {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
                                        for n, p in model.named_parameters()}


def _rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_native_step_is_as_close_to_fp32_as_bf16_autocast():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator(device="cuda:0").manual_seed(1)
        x = torch.randn(8, 1, 96, 96, device="cuda:0", generator=g)
        y = x + 0.3 * torch.randn(8, 1, 96, 96, device="cuda:0", generator=g)
        p_ref, l_ref, g_ref = _grads(_build(), x, y, native=False)
        p_ac, l_ac, g_ac = _grads(_build(), x, y, native=False, autocast=True)
        model = _build()
        p_nat, l_nat, g_nat = _grads(model, x, y, native=True)
        assert "_native_train_engine" in model.__dict__            # the CUDA path ran, not the module graph
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert abs(l_nat - l_ref) <= 1e-3 * abs(l_ref)
    assert _rel(p_nat, p_ref) <= max(6e-2, 1.2 * _rel(p_ac, p_ref))
    names = [n for n in g_ref if g_ref[n].norm() > 1e-6]
    err_nat = np.array([_rel(g_nat[n], g_ref[n]) for n in names])
    err_ac = np.array([_rel(g_ac[n], g_ref[n]) for n in names])
    assert np.median(err_nat) <= max(0.25, 1.2 * np.median(err_ac)), (np.median(err_nat), np.median(err_ac))
    assert err_nat.max() <= max(0.8, 1.5 * err_ac.max())
    # exactly-zero gradient for conv biases that feed a BatchNorm (documented deviation); torch has ~1e-9 noise there
    for n, gr in g_nat.items():
        if n.endswith("double_conv.0.bias") or n.endswith("double_conv.3.bias"):
            assert float(gr.abs().max()) == 0.0 and float(g_ref[n].abs().max()) < 1e-5


def test_backward_stat_fusion_matches_separate_reduction():
    """The optional BatchNorm-backward fusion into the data-gradient epilogues (engine.fuse_bwd_stats) and the forward
    statistics fusion (engine.fuse_stats) change where sums are computed, not what is computed."""
    g = torch.Generator(device="cuda:0").manual_seed(4)
    x = torch.randn(2, 1, 64, 64, device="cuda:0", generator=g)
    y = x + 0.3 * torch.randn(2, 1, 64, 64, device="cuda:0", generator=g)
    results = []
    from im2im_uq_b200.models.unet_train import UNetTrainEngine
    for fwd, bwd in ((True, False), (True, True), (False, False)):
        model = _build()
        eng = UNetTrainEngine(model)
        eng.fuse_stats, eng.fuse_bwd_stats = fwd, bwd
        model.__dict__["_native_train_engine"] = eng
        results.append(_grads(model, x, y, native=True))
    (p0, l0, g0), (p1, l1, g1), (p2, l2, g2) = results
    assert abs(l0 - l1) <= 2e-4 * abs(l0) and abs(l0 - l2) <= 2e-4 * abs(l0)   # statistics summed in a different order
    for n in g0:
        if g0[n].norm() > 1e-6:
            # Sums taken in a different pass differ in the last bits; one ulp in a BatchNorm scale re-rolls the bf16 rounding
            # of that layer's activations, so early layers see a fresh realisation of the bf16 noise (same level as
            # native-vs-fp32); the layers next to the loss agree tightly.  The sums themselves are checked to 2e-5 in
            # test_train_kernels_gpu.py.
            tol = 3e-2 if n.startswith("last_layer.") else 0.5
            assert _rel(g1[n], g0[n]) <= tol and _rel(g2[n], g0[n]) <= tol, (n, _rel(g1[n], g0[n]), _rel(g2[n], g0[n]))


def test_pool_fusion_is_a_pure_refactoring():
    """engine.fuse_pool: the same y / pooled / dy values element for element, so the loss is bit-identical and gradients agree
    to the order of the fp32 sums (also exercised with odd sizes, where the engine falls back to the separate kernels)."""
    from im2im_uq_b200.models.unet_train import UNetTrainEngine
    g = torch.Generator(device="cuda:0").manual_seed(9)
    for side in (64, 50):                      # 50 -> 25 -> 12: odd sizes down the path
        x = torch.randn(2, 1, side, side, device="cuda:0", generator=g)
        y = x + 0.3 * torch.randn(2, 1, side, side, device="cuda:0", generator=g)
        res = []
        for fuse in (True, False):
            model = _build()
            eng = UNetTrainEngine(model)
            eng.fuse_pool = fuse
            model.__dict__["_native_train_engine"] = eng
            res.append(_grads(model, x, y, native=True))
        (p0, l0, g0), (p1, l1, g1) = res
        # per element the fused kernels are bit-identical to the separate ones (test_train_kernels_gpu.py); two runs of the
        # whole network differ in the last bits anyway (BatchNorm statistics and weight gradients are fp32 atomic sums)
        # (one ulp in a BatchNorm scale re-rolls that layer's bf16 rounding: two runs agree to the bf16 noise level, ~2e-2)
        assert _rel(p0, p1) <= 4e-2 and abs(l0 - l1) <= 2e-3 * abs(l0)
        for n in g0:
            if g0[n].norm() > 1e-6:
                tol = 3e-2 if n.startswith("last_layer.") else 0.5      # see test_backward_stat_fusion_...: sums' last bits
                assert _rel(g1[n], g0[n]) <= tol, (side, n, _rel(g1[n], g0[n]))


MSE_PARAMS = dict(PARAMS, q_lo_weight=0.0, q_hi_weight=0.0, mse_weight=1.0)


def _build_mse():
    from core.models.add_uncertainty import add_uncertainty
    from core.models.trunks.unet import UNet
    torch.manual_seed(0)
    return add_uncertainty(UNet(1, 1), MSE_PARAMS).to("cuda:0").train()


def test_per_layer_gradients_with_a_smooth_loss():
    """With the (smooth) MSE part of the reference's loss only, the gradient is not sign-valued, so EVERY layer is checked
    against fp32 autograd through the same modules: relative L2 error of each parameter's gradient <= max(6e-2, 1.5 x the
    error torch's own bf16 autocast makes on that very layer) - bf16 rounding through 23 BatchNorm layers leaves ~0.2 on
    the first layers for any bf16 implementation - and the gradient norms agree within 15 %.  A wrong gradient in any
    single layer fails this (the pinball test above only bounds the median)."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator(device="cuda:0").manual_seed(7)
        x = torch.randn(8, 1, 96, 96, device="cuda:0", generator=g)
        y = x + 0.3 * torch.randn(8, 1, 96, 96, device="cuda:0", generator=g)
        p_ref, l_ref, g_ref = _grads(_build_mse(), x, y, native=False)
        p_ac, l_ac, g_ac = _grads(_build_mse(), x, y, native=False, autocast=True)
        m_nat = _build_mse()
        p_nat, l_nat, g_nat = _grads(m_nat, x, y, native=True)
        assert "_native_train_engine" in m_nat.__dict__
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert abs(l_nat - l_ref) <= 2e-3 * abs(l_ref)
    bad = {}
    for n in g_ref:
        if g_ref[n].norm() < 1e-7 or n.endswith("double_conv.0.bias") or n.endswith("double_conv.3.bias"):
            continue
        r, r_ac = _rel(g_nat[n], g_ref[n]), _rel(g_ac[n], g_ref[n])
        ratio = (g_nat[n].norm() / g_ref[n].norm()).item()
        if r > max(6e-2, 1.5 * r_ac) or not (0.85 <= ratio <= 1.18):
            bad[n] = (r, r_ac, ratio)
    assert not bad, bad
    # the layers next to the loss see almost no accumulated rounding: held to a tight absolute bound
    for n in g_ref:
        if n.startswith("last_layer.") and n.endswith("weight"):
            assert _rel(g_nat[n], g_ref[n]) <= 3e-2, (n, _rel(g_nat[n], g_ref[n]))


def test_data_parallel_arithmetic_on_one_gpu():
    """nn.DataParallel's arithmetic (train.py:22-27,112-115: per-replica BatchNorm statistics, loss over the gathered batch,
    summed gradients) against two native 'ranks' run one after the other on this GPU and averaged - the same comparison the
    2-GPU worker (tests/dp_train_worker.py) makes across real processes."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("dp_train_worker", os.path.join(os.path.dirname(__file__), "dp_train_worker.py"))
    w = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(w)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    world, B = 2, 16
    x = torch.randn(B, 1, 96, 96, generator=g).to(dev)
    y = (x.cpu() + 0.3 * torch.randn(B, 1, 96, 96, generator=g)).to(dev)
    ref, ref_loss = w.dataparallel_reference_grads(x, y, world, dev)
    ref_ac, _ = w.dataparallel_reference_grads(x, y, world, dev, autocast=True)
    total, losses = None, []
    for r in range(world):
        m = w.build(dev)
        m.zero_grad(set_to_none=True)
        loss = m.loss_fn(m(x.chunk(world)[r]), y.chunk(world)[r])
        loss.backward()
        losses.append(float(loss))
        gr = {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for n, p in m.named_parameters()}
        total = gr if total is None else {n: total[n] + gr[n] for n in gr}
    assert abs(sum(losses) / world - ref_loss) <= 2e-3 * abs(ref_loss)
    w.check_against_dataparallel({n: total[n] / world for n in total}, ref, ref_ac)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_data_parallel_step():
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29621", os.path.join(root, "tests", "dp_train_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "DP_TRAIN_OK" in res.stdout


def test_loss_trajectory_tracks_torch_adam():
    from im2im_uq_b200.models.unet_train import FusedAdam
    g = torch.Generator(device="cuda:0").manual_seed(2)
    x = torch.randn(4, 1, 64, 64, device="cuda:0", generator=g)
    y = x + 0.3 * torch.randn(4, 1, 64, 64, device="cuda:0", generator=g)
    m_ref, m_nat = _build(), _build()
    m_ref.use_native_training = False
    opt_ref = torch.optim.Adam(m_ref.parameters(), lr=1e-3)
    opt_nat = FusedAdam(m_nat.parameters(), lr=1e-3)
    for it in range(6):
        losses = []
        for model, opt in ((m_ref, opt_ref), (m_nat, opt_nat)):
            opt.zero_grad()
            loss = model.loss_fn(model(x), y)   # the reference's loop: train.py:152-153
            loss.backward()                     # train.py:160
            opt.step()                          # train.py:162
            losses.append(loss.item())
        assert abs(losses[1] - losses[0]) <= 3e-2 * abs(losses[0]), (it, losses)
    assert losses[1] < 0.5 * 2.1                # and it actually learns
    # BatchNorm running statistics advanced like torch's
    for (n1, b1), (n2, b2) in zip(m_ref.named_buffers(), m_nat.named_buffers()):
        # exact parity of the update rule is in test_train_kernels_gpu; here the bf16 and fp32 trajectories have drifted
        # apart for 6 optimizer steps, so only the variances (positive, O(1) scale) are compared, loosely
        if b1 is not None and "running_var" in n1:
            assert _rel(b2, b1) <= 0.3, (n1, _rel(b2, b1))
        if b1 is not None and "num_batches" in n1:
            assert int(b1) == int(b2) == 6
