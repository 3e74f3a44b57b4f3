"""CPU: host-side surface of the non-quantile heads - names, state dicts, loss formulas against reference KATs."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from im2im_uq_b200 import _lib
from im2im_uq_b200.models import heads
from im2im_uq_b200.models.add_uncertainty import add_uncertainty
from im2im_uq_b200.models.unet import UNet

PARAMS = dict(q_lo=0.05, q_hi=0.95, q_lo_weight=1.0, q_hi_weight=1.0, mse_weight=1.0, beta=0.1, num_softmax=50)
STATE_KEYS = {
    "gaussian": ["mean", "variance"], "residual_magnitude": ["prediction", "residual_magnitude"],
    "residual_magnitude_l1": ["prediction", "residual_magnitude"], "quantiles_l1": ["lower", "prediction", "upper"],
    "inn": ["lower", "prediction", "upper"], "softmax": ["output_layers.0"],
}


@pytest.mark.parametrize("head", sorted(STATE_KEYS))
def test_head_surface(head):
    torch.manual_seed(0)
    model = add_uncertainty(UNet(1, 1), dict(PARAMS, uncertainty_type=head))
    keys = [k.rsplit(".", 1)[0] for k in model.last_layer.state_dict() if k.endswith(".weight")]
    assert keys == STATE_KEYS[head]                                   # reference attribute names, reference order
    y = model(torch.randn(2, 1, 32, 32))
    planes = {"softmax": 50}.get(head, len(STATE_KEYS[head]))
    assert tuple(y.shape) == (2, planes, 1, 32, 32)
    if head in ("gaussian", "residual_magnitude", "residual_magnitude_l1"):
        assert (y[:, 1] >= 0).all()
    fn = model.in_nested_sets_from_output_fn
    assert fn.im2im_head_kind in (_lib.IM2IM_HEAD_QUANTILES, _lib.IM2IM_HEAD_RESIDUAL, _lib.IM2IM_HEAD_GAUSSIAN,
                                  _lib.IM2IM_HEAD_SOFTMAX_SETS)
    with pytest.raises(Exception, match="You have to specify lambda"):
        model.nested_sets_from_output(y)
    with pytest.raises(_lib.Im2ImError):                              # no CPU path behind the set functions
        model.nested_sets_from_output(y.detach(), 1.0)


def test_unknown_uncertainty_type_raises():
    with pytest.raises(NotImplementedError):
        add_uncertainty(UNet(1, 1), dict(PARAMS, uncertainty_type="bogus"))


def test_core_shim_exports_reference_names():
    import core.models.finallayers.gaussian_layer as gl
    import core.models.finallayers.inn_layer as il
    import core.models.finallayers.quantile_l1_layer as ql
    import core.models.finallayers.residual_magnitude_l1_layer as rl1
    import core.models.finallayers.residual_magnitude_layer as rl
    import core.models.finallayers.softmax_layer as sl
    from core.models.losses.inn import INNLoss
    assert gl.GaussianRegressionLayer is heads.GaussianRegressionLayer and gl.gaussian_regression_loss_fn
    assert rl.ResidualMagnitudeLayer and rl.residual_magnitude_nested_sets_from_output
    assert rl1.ResidualMagnitudeL1Layer and rl1.residual_magnitude_l1_loss_fn
    assert ql.QuantileRegressionL1Layer and ql.quantile_regression_l1_nested_sets_from_output
    assert il.INNLayer and il.inn_loss_fn and sl.SoftmaxLayer and sl.softmax_nested_sets_from_output
    with pytest.raises(AssertionError):
        INNLoss(beta=-1)


def test_head_loss_formulas_match_reference_kats_on_cpu():
    """The torch formulas used for CPU tensors equal the reference's losses (value and gradient)."""
    kats = np.load(os.path.join(GOLDEN, "head_loss_kats.npz"))
    fns = dict(gaussian=heads.gaussian_regression_loss_fn, residual_magnitude=heads.residual_magnitude_loss_fn,
               residual_magnitude_l1=heads.residual_magnitude_l1_loss_fn, quantiles_l1=heads.quantile_regression_l1_loss_fn,
               inn=heads.inn_loss_fn, softmax=heads.softmax_loss_fn)
    for head, fn in fns.items():
        for tag in ("a", "b"):
            key = f"{head}_{tag}"
            params = json.loads(str(kats[key + "_params"]))
            pred = torch.from_numpy(kats[key + "_pred"]).requires_grad_(True)
            loss = fn(pred, torch.from_numpy(kats[key + "_target"]), params)
            loss.backward()
            np.testing.assert_allclose(loss.item(), float(kats[key + "_loss"]), rtol=1e-6)
            np.testing.assert_allclose(pred.grad.numpy(), kats[key + "_grad"], rtol=1e-5, atol=1e-9)
