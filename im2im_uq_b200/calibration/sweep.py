"""Host replay of the reference's lambda-sweep stopping rule on top of one-pass miss counts.

The reference (core/calibration/calibrate_model.py:130-145) walks the lambda grid from the top, and at every step
    losses = per-image fp32 miss fractions at lam - dlambda          (:135, a full pass over the data)
    Rhat = losses.mean()                                             (:137, fp32 torch.mean on CPU)
    RhatPlus = HB_mu_plus(Rhat.item(), N, delta)                     (:138, float64 scipy)
    if Rhat >= alpha or RhatPlus > alpha: lhat = lam; break          (:140-144)
Our kernel delivers ALL columns at once as exact integers, so the walk needs no data pass.  To return the
reference's answer bit for bit the decision is still taken with the reference's own expression
(:func:`exact_stop_condition`) - but only on the few columns where the outcome is not already certain:

  * per-column totals T_j (exact int64) give the exact risk R_j = T_j / (N*px);
  * the reference's fp32 Rhat differs from R_j by at most gamma*R_j, gamma = (N+8)*2^-24 (per-image rounding of
    count/px, any summation order over N non-negative terms, final division);
  * HB_mu_plus is non-decreasing in muhat > 0, so ``HB > alpha`` is decided by comparing against the cached bracket
    of its level set (bounds.hb_stop_bracket) whenever R_j is outside the guard band;
  * columns inside the band, and columns with T_j == 0 (where the reference's bound returns 1.0 through its
    exception path), are replayed exactly.

This module is pure host logic (numpy/torch CPU); it is exercised on CPU with counts from the oracle and on GPU with
counts from the CUDA kernel.
"""
import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from .bounds import HB_mu_plus, hb_stop_bracket

_U = 2.0 ** -24  # fp32 unit roundoff
_HB_SLACK = 1e-9  # absolute guard around the cached HB level set (brentq tolerance is 2e-12)


def exact_stop_condition(losses: torch.Tensor, n: int, alpha: float, delta: float, verbose_lam=None) -> bool:
    """The reference's per-step decision, verbatim in meaning (calibrate_model.py:137-140).

    ``losses`` must be the (N,) fp32 CPU tensor of per-image losses, in image order, so that torch's CPU mean sees
    the same values in the same order as the reference's ``torch.cat(losses)``.
    """
    assert losses.device.type == "cpu" and losses.dtype == torch.float32
    Rhat = losses.mean()
    RhatPlus = HB_mu_plus(Rhat.item(), n, delta)
    if verbose_lam is not None:
        print(f"\rLambda: {verbose_lam:.4f}  |  Rhat: {Rhat:.4f}  |  RhatPlus: {RhatPlus:.4f}  ", end='')
    return bool(Rhat >= alpha or RhatPlus > alpha)


def screening_constants(n: int, px: int, alpha: float, delta: float):
    """(N*px, gamma, alpha32, r_lo, r_hi, slack): the numbers both the host screen (_classify) and the device screen
    (im2im_rcps_decide) are built from."""
    gamma = (n + 8) * _U * 1.01
    alpha32 = float(np.float32(alpha))  # `Rhat >= alpha` compares in fp32 (0-dim fp32 tensor vs python float)
    r_lo, r_hi = hb_stop_bracket(int(n), float(alpha), float(delta))
    return float(n) * float(px), gamma, alpha32, r_lo, r_hi, _HB_SLACK


def _classify(totals: np.ndarray, n: int, px: int, alpha: float, delta: float):
    """Per column: +1 certainly stops, -1 certainly continues, 0 must be replayed exactly."""
    n_px, gamma, alpha32, r_lo, r_hi, _ = screening_constants(n, px, alpha, delta)
    R = totals.astype(np.float64) / n_px
    lo_R, hi_R = R * (1.0 - gamma), R * (1.0 + gamma)  # the reference's fp32 Rhat lies in [lo_R, hi_R]
    verdict = np.zeros(R.shape, dtype=np.int8)
    # certainly true: either clause certainly true
    sure_true = lo_R >= alpha32 * (1.0 + 1e-6)
    if math.isfinite(r_hi):
        sure_true |= lo_R > r_hi + _HB_SLACK
    # certainly false: both clauses certainly false
    sure_false = hi_R < alpha32 * (1.0 - 1e-6)
    if math.isfinite(r_lo):
        sure_false &= hi_R < r_lo - _HB_SLACK
    verdict[sure_true] = 1
    verdict[sure_false & ~sure_true] = -1
    verdict[totals == 0] = 0  # HB_mu_plus(0) takes the reference's exception path; never guess it
    return verdict


def find_stop_index(totals: Sequence[int], n: int, px: int, alpha: float, delta: float,
                    column_losses: Callable[[int], torch.Tensor], monotone: bool = True,
                    verbose_lambdas: Optional[torch.Tensor] = None, stats: Optional[dict] = None) -> int:
    """Index j of the column at which the reference's reverse scan stops, or -1 if it runs off the grid.

    totals[j]        exact miss total of column j over all N images (any integer dtype)
    column_losses(j) -> (N,) fp32 CPU tensor of per-image losses of column j (called only for replayed columns)
    monotone         totals are non-decreasing along the scan (ascending lambda grid).  When False every column is
                     replayed exactly, in scan order, like the reference.
    """
    totals = np.asarray(totals, dtype=np.int64)
    L = totals.shape[0]
    replayed = 0
    scan = range(L - 1, -1, -1)
    if monotone and L > 1 and np.any(np.diff(totals) > 0):
        monotone = False  # not the shape the screening argument needs (e.g. a descending grid): replay everything
    if not monotone:
        verdict = np.zeros(L, dtype=np.int8)
    else:
        verdict = _classify(totals, n, px, alpha, delta)
    stop = -1
    for j in scan:
        v = verdict[j]
        if v < 0:
            continue
        if v > 0:
            stop = j
            break
        replayed += 1
        lam = None if verbose_lambdas is None else float(verbose_lambdas[j])
        if exact_stop_condition(column_losses(j), n, alpha, delta, verbose_lam=lam):
            stop = j
            break
    if stats is not None:
        stats["replayed_columns"] = replayed
        stats["screened"] = bool(monotone)
    return stop


def lambda_grid(config: dict):
    """(lambdas, dlambda, lam_prime, default_lhat) exactly as calibrate_model.py:97-100,130-131,135 builds them."""
    if config["uncertainty_type"] == "softmax":
        lambdas = torch.linspace(config['minimum_lambda_softmax'], config['maximum_lambda_softmax'],
                                 config['num_lambdas'])
    else:
        lambdas = torch.linspace(config['minimum_lambda'], config['maximum_lambda'], config['num_lambdas'])
    dlambda = lambdas[1] - lambdas[0]  # IndexError for a one-point grid, like the reference
    lam_prime = lambdas - dlambda      # elementwise fp32: identical to the per-step `lam - dlambda`
    default_lhat = lambdas[-1] + dlambda - 1e-9
    return lambdas, dlambda, lam_prime, default_lhat


def visited_mask(lambdas: torch.Tensor, stop: int) -> torch.Tensor:
    """Columns the reference's scan wrote (calibrate_model.py:136 writes every column whose lambda equals lam)."""
    L = lambdas.shape[0]
    if stop < 0:
        return torch.ones(L, dtype=torch.bool)
    seen = lambdas[stop:]
    return torch.isin(lambdas, seen)


def sweep_from_counts(counts: torch.Tensor, totals: torch.Tensor, px: int, config: dict, column_to_losses,
                      ascending: bool = True, group=None, verbose: bool = False, stats: Optional[dict] = None,
                      n_total: Optional[int] = None, device_decide=None, totals_already_reduced: bool = False):
    """Stopping rule + lambda-hat from this rank's integer miss counts; the multi-GPU exchange lives here.

    counts            (N_local, L) int32, rows = this rank's images in order, columns = the ORIGINAL grid order
    totals            (L,) int64 column sums of ``counts`` (this rank only); all-reduced in place when ``group`` is set
    column_to_losses  f(counts[:, j]) -> (N_local,) fp32 per-image losses (count/px as the device or host computes it)
    group             None for one process, else a torch.distributed process group: ranks hold contiguous shards of
                      the calibration set in rank order.  Collectives: ONE all_reduce of the int64 totals (8*L bytes),
                      one all_gather of the shard sizes (skipped when ``n_total`` is given and no column has to be
                      replayed), and an all_gather of N fp32 values per replayed column.
    n_total           total number of images over all ranks, if the caller knows it
    device_decide     optional f(totals, n_total) -> (stop, decided) running the screen on the device
                      (im2im_rcps_decide); when it decides, the host never sees the totals.
    Returns (lhat 0-dim fp32 CPU tensor, stop index or -1, visited bool mask (L,)).
    """
    lambdas, dlambda, lam_prime, default_lhat = lambda_grid(config)
    n_local = counts.shape[0]
    sizes = None

    def exchange_sizes():
        nonlocal sizes, px
        import torch.distributed as dist
        meta = torch.tensor([n_local, px], dtype=torch.int64, device=totals.device)
        gathered = [torch.zeros_like(meta) for _ in range(dist.get_world_size(group))]
        dist.all_gather(gathered, meta, group=group)
        sizes = [int(t[0]) for t in gathered]
        px = max(int(t[1]) for t in gathered)
        return sum(sizes)

    if group is not None:
        import torch.distributed as dist
        if not totals_already_reduced:
            dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)
        if n_total is None:
            n_total = exchange_sizes()
    elif n_total is None:
        n_total = n_local
    if n_total == 0:
        raise ValueError("empty calibration set")
    if device_decide is not None and ascending:
        stop, decided = device_decide(totals, n_total)
        if decided:
            if stats is not None:
                stats["replayed_columns"] = 0
                stats["screened"] = True
                stats["decided_on_device"] = True
            lhat = lambdas[stop] if stop >= 0 else default_lhat
            return lhat, stop, visited_mask(lambdas, stop)
    if group is not None and sizes is None:
        exchange_sizes()
    totals_host = totals.cpu().numpy()

    def column_losses(j: int) -> torch.Tensor:
        col = column_to_losses(counts[:, j])
        if sizes is not None:
            import torch.distributed as dist
            pad = torch.zeros(max(sizes), dtype=torch.float32, device=col.device)
            pad[:n_local] = col
            parts = [torch.zeros_like(pad) for _ in sizes]
            dist.all_gather(parts, pad, group=group)
            col = torch.cat([p[:k] for p, k in zip(parts, sizes)])
        return col.cpu()

    stop = find_stop_index(totals_host, n_total, px, config['alpha'], config['delta'], column_losses,
                           monotone=ascending, verbose_lambdas=lambdas if verbose else None, stats=stats)
    lhat = lambdas[stop] if stop >= 0 else default_lhat
    return lhat, stop, visited_mask(lambdas, stop)
