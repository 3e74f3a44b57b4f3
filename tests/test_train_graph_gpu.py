"""GPU (-m gpu): the training iteration captured into a CUDA graph (GraphedTrainStep) follows the eager native loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu

PARAMS = dict(uncertainty_type="quantiles", q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0)


def _build():
    from core.models.add_uncertainty import add_uncertainty
    from core.models.trunks.unet import UNet
    torch.manual_seed(0)
    return add_uncertainty(UNet(1, 1), PARAMS).to("cuda:0").train()


def test_graphed_step_tracks_eager_step():
    from im2im_uq_b200 import _lib
    from im2im_uq_b200.models.unet_train import FusedAdam, GraphedTrainStep
    g = torch.Generator(device="cuda:0").manual_seed(2)
    xs = [torch.randn(4, 1, 64, 64, device="cuda:0", generator=g) for _ in range(3)]
    ys = [x + 0.3 * torch.randn(4, 1, 64, 64, device="cuda:0", generator=g) for x in xs]
    warm = 2
    # eager native loop over the sequence; the graph's warm-up steps on batch 0 are rolled back before capture (parameters,
    # Adam state, BatchNorm buffers), so the first replay is the first step of the trajectory
    m_e = _build()
    o_e = FusedAdam(m_e.parameters(), lr=1e-3)
    eager = []
    seq = [1, 2, 0, 1, 2]
    for i in seq:
        o_e.zero_grad()
        loss = m_e.loss_fn(m_e(xs[i]), ys[i])
        loss.backward()
        o_e.step()
        eager.append(loss.item())
    m_g = _build()
    o_g = FusedAdam(m_g.parameters(), lr=1e-3)
    step = GraphedTrainStep(m_g, o_g, xs[0], ys[0], warmup=warm)
    assert step.kernels_per_replay > 100
    before = _lib.launch_count()
    graph_losses = [step(xs[i].cpu().pin_memory(), ys[i].cpu().pin_memory()).item() for i in [1, 2, 0, 1, 2]]
    assert _lib.launch_count() == before         # replays launch nothing from the host side
    # wgrad uses fp32 atomics (order varies run to run), so trajectories agree to rounding, not bit for bit
    assert step.warmup_steps == 0
    for a, b in zip(graph_losses, eager):
        assert abs(a - b) <= 2e-2 * abs(b), (graph_losses, eager)
    assert graph_losses[-1] < graph_losses[0]
    # device-side step counter advanced once per replay (the warm-up was rolled back)
    assert int(o_g.state_dev[0].item()) == 5
    for (n1, b1), (n2, b2) in zip(m_e.named_buffers(), m_g.named_buffers()):
        if b1 is not None and "num_batches" in n1:
            assert int(b1) == int(b2) == len(seq)


def test_eval_engine_sees_graph_replays_and_fused_adam_steps():
    """train (graph replay) -> eval -> more replays -> eval: the cached inference engine must re-fold the weights every
    time, although FusedAdam / bn_finalize / cudaGraphLaunch never bump a tensor._version."""
    from im2im_uq_b200.models.unet_train import FusedAdam, GraphedTrainStep
    g = torch.Generator(device="cuda:0").manual_seed(3)
    x = torch.randn(4, 1, 64, 64, device="cuda:0", generator=g)
    y = x + 0.3 * torch.randn(4, 1, 64, 64, device="cuda:0", generator=g)
    model = _build()
    opt = FusedAdam(model.parameters(), lr=1e-2)
    step = GraphedTrainStep(model, opt, x, y, warmup=1)

    def eval_pair():
        model.eval()
        with torch.no_grad():
            native = model(x)
            assert "_native_engine" in model.__dict__
            model.use_native_inference = False
            ref = model(x)                        # torch module graph on the CURRENT parameters / running statistics
            model.use_native_inference = True
        model.train()
        return native, ref

    outs = []
    for _ in range(3):
        for _ in range(2):
            step(x, y)
        native, ref = eval_pair()
        assert ((native - ref).norm() / ref.norm()).item() <= 2e-2
        outs.append(native)
    # the model really moved between the evaluations (lr 1e-2): a stale engine would have returned the same tensor
    assert ((outs[1] - outs[0]).norm() / outs[0].norm()).item() > 5e-2
    assert ((outs[2] - outs[1]).norm() / outs[1].norm()).item() > 1e-2
    # same for the eager native loop with FusedAdam
    for _ in range(2):
        opt.zero_grad()
        model.loss_fn(model(x), y).backward()
        opt.step()
    native, ref = eval_pair()
    assert ((native - ref).norm() / ref.norm()).item() <= 2e-2
    step.close()
