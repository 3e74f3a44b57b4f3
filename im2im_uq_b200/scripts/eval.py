"""Evaluation-side callers of the hot path - mirror of the reference's ``core/scripts/eval.py``:

    get_loss_table(model, dataset, config) -> (N_val, L) fp32 CPU table at lambdas[j]       reference :84-126
    eval_set_metrics(model, dataset, config) -> (risk, sizes, spearman, stratified risks,    reference :129-157
                                                 mse, spatial miscoverage)

plus the loss-table wire format the reference's router writes and its plot scripts read
(core/scripts/router.py:138: ``torch.save(torch.cat((calib_loss_table, val_loss_table), dim=0), ...)``;
experiments/*/plot.py:126-139 ``plot_risks``: ``num_trials`` random splits through ``evaluate_from_loss_table``).

Underneath: the model's outputs stay in HBM (native UNet forward), the dense table is ONE pass of
``im2im_rcps_miss_counts`` over the un-shifted grid (the reference runs L x ceil(N/4) tiny launches, :118-125), the
metrics are the fused kernels of ``get_rcps_metrics_from_outputs``.  ``get_images`` / ``eval_net`` are wandb plumbing
and out of scope.
"""
import torch

from .. import rcps
from ..calibration import calibrate_model as cm
from ..calibration import sweep


def _reset(dataset):
    try:
        dataset.reset()
    except Exception:
        print("dataset is map-style (not resettable)")


def get_loss_table(model, dataset, config):
    _reset(dataset)
    with torch.no_grad():
        lambdas = sweep.lambda_grid(config)[0]             # same keys as the reference, incl. the softmax grid (:90-93)
        model.eval()
        device = cm._cuda_device(config['device'])
        cm.get_rcps_loss_fn(config)
        model = model.to(device)
        ascending = bool((lambdas[1:] >= lambdas[:-1]).all()) if lambdas.numel() > 1 else True
        lam_sorted, order = (lambdas, None) if ascending else torch.sort(lambdas)
        if cm.streaming_applicable(model, dataset, config) and len(dataset) > 0:
            # batch by batch, the (N, 3, C, H, W) tensor of eval.py:100-112 is never built (quantile head: the head
            # convolution's epilogue books the ranks itself)
            print("GET LOSS TABLE FROM OUTPUTS")
            counts, _, px = cm.stream_miss_counts(model, dataset, config, device, lam_sorted.to(device))
        else:
            outputs, labels = cm.collect_outputs(model, dataset, config, device)
            kind, scores = cm._head_scores(model, outputs, device)
            print("GET LOSS TABLE FROM OUTPUTS")
            counts, _ = rcps.miss_counts(scores, labels, lam_sorted.to(device), head=kind)
            px = max(labels[0].numel(), 1) if labels.shape[0] else 1
        if order is not None:
            inv = torch.empty_like(order)
            inv[order] = torch.arange(order.numel())
            counts = counts[:, inv.to(device)].contiguous()
        table = rcps.loss_table(counts, px)
        print("DONE!")
        return table.cpu()


def eval_set_metrics(model, dataset, config):
    _reset(dataset)
    with torch.no_grad():
        model.eval()
        device = cm._cuda_device(config['device'])
        rcps_loss_fn = cm.get_rcps_loss_fn(config)
        model = model.to(device)
        outputs, labels = cm.collect_outputs(model, dataset, config, device)
        print("GET RCPS METRICS FROM OUTPUTS")
        losses, sizes, spearman, stratified_risks, mse, spatial_miscoverage = cm.get_rcps_metrics_from_outputs(
            model, (outputs, labels), rcps_loss_fn, device)
        print("DONE!")
        return losses.mean(), sizes, spearman, stratified_risks, mse, spatial_miscoverage


# ------------------------------------------------------------------------------------------------ loss-table files
def loss_table_filename(config) -> str:
    """File name of router.py:138 (without the output directory)."""
    return (f"loss_table_{config['dataset']}_{config['uncertainty_type']}_{config['batch_size']}_{config['lr']}_"
            f"{config['input_normalization']}_" + str(config['output_normalization']).replace('.', '_') + ".pth")


def save_loss_tables(calib_loss_table: torch.Tensor, val_loss_table: torch.Tensor, path: str) -> torch.Tensor:
    """The reference's wire format: ONE (N_calib + N_val, L) fp32 CPU tensor, calibration rows first (router.py:138)."""
    table = torch.cat((calib_loss_table.detach().float().cpu(), val_loss_table.detach().float().cpu()), dim=0)
    torch.save(table, path)
    return table


def load_loss_table(path: str) -> torch.Tensor:
    table = torch.load(path, map_location="cpu")
    if not torch.is_tensor(table) or table.dim() != 2 or table.dtype != torch.float32:
        raise ValueError(f"{path}: not a reference loss table (2-D fp32 tensor)")
    return table


def evaluate_loss_table_trials(loss_table: torch.Tensor, n: int, alpha: float, delta: float, num_trials: int = 100):
    """The trial loop of plot_risks (experiments/*/plot.py:133-136): ``num_trials`` calls of evaluate_from_loss_table,
    consuming the global torch RNG exactly like the reference (one randperm per trial).  Each trial's Hoeffding-Bentkus
    scan is screened through the cached level set (bounds.hb_stop_bracket) instead of one brentq solve per column."""
    risks = torch.zeros((num_trials,))
    for trial in range(num_trials):
        risks[trial] = cm.evaluate_from_loss_table(loss_table, n, alpha, delta)
    return risks
