"""ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement of the reference's RCPS calibration path (quantile head).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module, and only as the checker or as the timed CPU baseline.  The product package (``im2im_uq_b200``) never
imports it and fails loudly when its CUDA library is missing.

Parity status: PINNED against ``tests/golden/rcps_*.npz`` and ``tests/golden/hb_mu_plus_kats.json``, produced by
running the unmodified reference in the authoring container (``tests/golden/make_golden.py``).  The reference
ships no golden vectors of its own.

Two independent restatements of the per-pixel chain are provided and cross-checked in tests:
  * ``librcps_oracle.so`` (``rcps_oracle.c``; plain C, fp32, -ffp-contract=off) - fast enough for 10^8 pixel-steps
  * ``np_*`` functions (numpy float32 array ops, one rounding per op like ATen) - the most literal transcription

Reference lines followed (relative to the reference root):
  core/models/finallayers/quantile_layer.py:34-44, core/models/add_uncertainty.py:33-38,
  core/calibration/calibrate_model.py:76-80 (loss), :130-145 (sweep), core/calibration/bounds.py:6-29 (bound).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_F32P = ctypes.POINTER(ctypes.c_float)
_I32P = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """Compile rcps_oracle.c with the recipe in oracle/Makefile.  Building the checker is not using it."""
    so = os.path.join(_HERE, "librcps_oracle.so")
    src = os.path.join(_HERE, "rcps_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "librcps_oracle.so"], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        i64, f32 = ctypes.c_int64, ctypes.c_float
        lib.oracle_quantile_miss_counts.argtypes = [_F32P] * 4 + [i64] * 6 + [f32, _I32P]
        lib.oracle_quantile_miss_table.argtypes = [_F32P] * 4 + [i64] * 6 + [_F32P, i64, _I32P]
        lib.oracle_quantile_nested_sets.argtypes = [_F32P] * 3 + [i64] * 5 + [f32, _F32P, _F32P]
        lib.oracle_quantile_miss_map.argtypes = [_F32P] * 4 + [i64] * 6 + [f32, _I32P]
        i32 = ctypes.c_int32
        lib.oracle_head_miss_table.argtypes = [i32] + [_F32P] * 4 + [i64] * 6 + [_F32P, i64, _I32P]
        lib.oracle_head_nested_sets.argtypes = [i32] + [_F32P] * 3 + [i64] * 5 + [f32, _F32P, _F32P]
        lib.oracle_head_miss_map.argtypes = [i32] + [_F32P] * 4 + [i64] * 6 + [f32, _I32P]
        lib.oracle_softmax_sets.argtypes = [_F32P, i64, i64, i64, _F32P]
        for fn in (lib.oracle_head_miss_table, lib.oracle_head_nested_sets, lib.oracle_head_miss_map,
                   lib.oracle_softmax_sets):
            fn.restype = None
        lib.oracle_num_threads.restype = ctypes.c_int
        for fn in (lib.oracle_quantile_miss_counts, lib.oracle_quantile_miss_table,
                   lib.oracle_quantile_nested_sets, lib.oracle_quantile_miss_map):
            fn.restype = None
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().oracle_num_threads())


def _planes(outputs: np.ndarray, labels: np.ndarray):
    """outputs (N,3,C,H,W) fp32, labels (N,C,H,W) fp32 -> contiguous arrays, pixel count and element strides."""
    outputs = np.ascontiguousarray(outputs, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.float32)
    n = outputs.shape[0]
    assert outputs.shape[1] == 3 and labels.shape[0] == n
    px = int(np.prod(outputs.shape[2:]))
    assert int(np.prod(labels.shape[1:])) == px
    return outputs, labels, n, px


def _ptr(a: np.ndarray, offset_elems: int = 0):
    return ctypes.cast(a.ctypes.data + 4 * offset_elems, _F32P)


def c_miss_counts(outputs, labels, lam: float) -> np.ndarray:
    """Integer miss count per image at ONE lambda (one full pass, as calibrate_model.py:135 does per step)."""
    outputs, labels, n, px = _planes(outputs, labels)
    counts = np.zeros(n, dtype=np.int32)
    _lib().oracle_quantile_miss_counts(_ptr(outputs, 0), _ptr(outputs, px), _ptr(outputs, 2 * px), _ptr(labels),
                                       n, px, 3 * px, 3 * px, 3 * px, px, np.float32(lam),
                                       counts.ctypes.data_as(_I32P))
    return counts


def c_miss_table(outputs, labels, lams) -> np.ndarray:
    """(N, L) integer miss counts, one pass per lambda."""
    outputs, labels, n, px = _planes(outputs, labels)
    lams = np.ascontiguousarray(lams, dtype=np.float32)
    counts = np.zeros((n, lams.shape[0]), dtype=np.int32)
    _lib().oracle_quantile_miss_table(_ptr(outputs, 0), _ptr(outputs, px), _ptr(outputs, 2 * px), _ptr(labels),
                                      n, px, 3 * px, 3 * px, 3 * px, px, _ptr(lams), lams.shape[0],
                                      counts.ctypes.data_as(_I32P))
    return counts


def c_nested_sets(outputs, lam: float):
    outputs = np.ascontiguousarray(outputs, dtype=np.float32)
    n = outputs.shape[0]
    px = int(np.prod(outputs.shape[2:]))
    lo = np.empty((n,) + outputs.shape[2:], dtype=np.float32)
    up = np.empty_like(lo)
    _lib().oracle_quantile_nested_sets(_ptr(outputs, 0), _ptr(outputs, px), _ptr(outputs, 2 * px), n, px,
                                       3 * px, 3 * px, 3 * px, np.float32(lam), _ptr(lo), _ptr(up))
    return lo, outputs[:, 1].copy(), up


def c_miss_map(outputs, labels, lam: float) -> np.ndarray:
    outputs, labels, n, px = _planes(outputs, labels)
    m = np.zeros(px, dtype=np.int32)
    _lib().oracle_quantile_miss_map(_ptr(outputs, 0), _ptr(outputs, px), _ptr(outputs, 2 * px), _ptr(labels),
                                    n, px, 3 * px, 3 * px, 3 * px, px, np.float32(lam), m.ctypes.data_as(_I32P))
    return m.reshape(outputs.shape[2:])


# ----------------------------------------------------------------------------- other heads (C restatement)
HEAD_KINDS = {"quantiles": 0, "quantiles_l1": 0, "inn": 0, "residual_magnitude": 1, "residual_magnitude_l1": 1,
              "gaussian": 2, "softmax_sets": 3}


def _head_planes(outputs: np.ndarray, head: str):
    """(a, p, b) plane offsets (in elements) and image stride for the head's output tensor (N, 2 or 3, ...)."""
    kind = HEAD_KINDS[head]
    outputs = np.ascontiguousarray(outputs, dtype=np.float32)
    n = outputs.shape[0]
    px = int(np.prod(outputs.shape[2:]))
    if kind in (0, 3):
        assert outputs.shape[1] == 3
        offs, stride = (0, px, 2 * px), 3 * px
    else:
        assert outputs.shape[1] == 2
        offs, stride = (0, 0, px), 2 * px   # a is ignored for the 2-plane heads
    return kind, outputs, n, px, offs, stride


def head_miss_table(outputs, labels, lams, head: str) -> np.ndarray:
    """(N, L) integer miss counts for any head, one pass per lambda."""
    kind, outputs, n, px, offs, stride = _head_planes(outputs, head)
    labels = np.ascontiguousarray(labels, dtype=np.float32)
    lams = np.ascontiguousarray(np.atleast_1d(lams), dtype=np.float32)
    counts = np.zeros((n, lams.shape[0]), dtype=np.int32)
    _lib().oracle_head_miss_table(kind, _ptr(outputs, offs[0]), _ptr(outputs, offs[1]), _ptr(outputs, offs[2]),
                                  _ptr(labels), n, px, stride, stride, stride, px, _ptr(lams), lams.shape[0],
                                  counts.ctypes.data_as(_I32P))
    return counts


def head_nested_sets(outputs, lam: float, head: str):
    kind, outputs, n, px, offs, stride = _head_planes(outputs, head)
    lo = np.empty((n,) + outputs.shape[2:], dtype=np.float32)
    up = np.empty_like(lo)
    _lib().oracle_head_nested_sets(kind, _ptr(outputs, offs[0]), _ptr(outputs, offs[1]), _ptr(outputs, offs[2]), n, px,
                                   stride, stride, stride, np.float32(lam), _ptr(lo), _ptr(up))
    pred = outputs[:, 1] if kind in (0, 3) else outputs[:, 0]
    return lo, pred.copy(), up


def head_miss_map(outputs, labels, lam: float, head: str) -> np.ndarray:
    kind, outputs, n, px, offs, stride = _head_planes(outputs, head)
    labels = np.ascontiguousarray(labels, dtype=np.float32)
    m = np.zeros(px, dtype=np.int32)
    _lib().oracle_head_miss_map(kind, _ptr(outputs, offs[0]), _ptr(outputs, offs[1]), _ptr(outputs, offs[2]),
                                _ptr(labels), n, px, stride, stride, stride, px, np.float32(lam),
                                m.ctypes.data_as(_I32P))
    return m.reshape(outputs.shape[2:])


def softmax_sets(logits: np.ndarray) -> np.ndarray:
    """softmax_layer.py:34-48: logits (N, K, ...) -> (N, 3, ...) = (lower quantile, prediction, upper quantile)."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    n, k = logits.shape[:2]
    inner = int(np.prod(logits.shape[2:]))
    sets = np.empty((n, 3) + logits.shape[2:], dtype=np.float32)
    _lib().oracle_softmax_sets(_ptr(logits), n, k, inner, _ptr(sets))
    return sets


def np_head_nested_sets(outputs: np.ndarray, lam, head: str):
    """Literal numpy transcription of the other heads' set functions + add_uncertainty.py:35-36."""
    kind = HEAD_KINDS[head]
    if kind == 0:
        return np_nested_sets(outputs, lam)
    eps, lam = np.float32(1e-6), np.float32(lam)
    out = np.asarray(outputs, dtype=np.float32)
    with np.errstate(all="ignore"):
        if kind == 3:
            a, p, b = out[:, 0], out[:, 1], out[:, 2]
            relu = lambda t: np.where(np.isnan(t), t, np.maximum(t, np.float32(0)))
            lower = p - relu(p - a) * lam
            upper = p + relu(b - p) * lam
        else:
            p = out[:, 0]
            w = np.sqrt(out[:, 1]) if kind == 2 else out[:, 1]
            upper = lam * w + p
            lower = -lam * w + p
        upper = np.maximum(upper, p + eps)
        lower = np.minimum(lower, p - eps)
    return lower, p, upper


# ----------------------------------------------------------------------------- numpy transcription
def np_nested_sets(outputs: np.ndarray, lam) -> tuple:
    """quantile_layer.py:39-42 then add_uncertainty.py:35-36; every numpy op rounds to fp32 like an ATen op."""
    eps = np.float32(1e-6)
    lam = np.float32(lam)
    out = np.array(outputs, dtype=np.float32, copy=True)
    with np.errstate(all="ignore"):
        l, p, u = out[:, 0], out[:, 1], out[:, 2]
        l = np.minimum(l, p - eps)          # np.minimum/maximum propagate NaN like torch.minimum/maximum
        u = np.maximum(u, p + eps)
        upper = lam * (u - p) + p
        lower = p - lam * (p - l)
        upper = np.maximum(upper, p + eps)
        lower = np.minimum(lower, p - eps)
    return lower, p, upper


def np_fraction_missed(sets, labels: np.ndarray) -> tuple:
    """calibrate_model.py:76-80 -> (per-image fp32 loss, per-image integer miss count)."""
    lower, _, upper = sets
    labels = np.asarray(labels, dtype=np.float32)
    with np.errstate(all="ignore"):
        misses = (lower > labels).astype(np.float32) + (upper < labels).astype(np.float32)
    misses[misses > 1.0] = 1.0
    n = misses.shape[0]
    counts = misses.reshape(n, -1).sum(axis=1, dtype=np.float64).astype(np.int32)
    px = np.float32(misses.reshape(n, -1).shape[1])
    return counts.astype(np.float32) / px, counts


def np_miss_counts(outputs, labels, lam) -> np.ndarray:
    return np_fraction_missed(np_nested_sets(outputs, lam), labels)[1]


# ----------------------------------------------------------------------------- bound + sweep restatement
def h1(y, mu):
    """bounds.py:6-7"""
    return y * np.log(y / mu) + (1 - y) * np.log((1 - y) / (1 - mu))


def hb_mu_plus(muhat, n, delta, maxiters=1000):
    """bounds.py:17-29 - Hoeffding-Bentkus upper confidence bound (scipy brentq + binom.cdf, like the reference)."""
    from scipy.optimize import brentq
    from scipy.stats import binom

    def tail(mu):
        hoeff = -n * h1(np.minimum(mu, muhat), mu)                                     # bounds.py:10-11
        bent = np.log(max(binom.cdf(np.floor(n * muhat), n, mu), 1e-10)) + 1           # bounds.py:13-14
        return min(hoeff, bent) - np.log(delta)

    if tail(1 - 1e-10) > 0:
        return 1
    try:
        return brentq(tail, muhat, 1 - 1e-10, maxiter=maxiters)
    except Exception:  # the reference's bare except: includes muhat == 0 (0*log 0 = nan)
        return 1.0


def calibrate_sweep(outputs, labels, lam_min, lam_max, num_lambdas, alpha, delta, miss_counts=c_miss_counts,
                    head=None):
    """calibrate_model.py:97-100,130-145 restated: reverse linear scan with early stop.

    Returns (lhat fp32 0-dim tensor, stop index or -1, (N,L) fp32 loss table with unvisited columns zero).
    torch is used for linspace / fp32 mean because the reference's numbers come from exactly those torch CPU ops.
    """
    import torch
    import warnings

    if head is not None:
        miss_counts = lambda o, l, lam: head_miss_table(o, l, [lam], head)[:, 0]  # noqa: E731
    lambdas = torch.linspace(lam_min, lam_max, num_lambdas)
    n = outputs.shape[0]
    px = int(np.prod(outputs.shape[2:]))
    dlambda = lambdas[1] - lambdas[0]
    lhat = lambdas[-1] + dlambda - 1e-9
    table = torch.zeros((n, num_lambdas))
    stop = -1
    for j in reversed(range(num_lambdas)):
        lam = lambdas[j]
        counts = miss_counts(outputs, labels, float(lam - dlambda))
        losses = torch.from_numpy(counts.astype(np.float32)) / float(px)
        table[:, np.where(lambdas.numpy() == lam.numpy())[0]] = losses[:, None]
        rhat = losses.mean()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rhat_plus = hb_mu_plus(rhat.item(), n, delta)
        if rhat >= alpha or rhat_plus > alpha:
            lhat, stop = lam, j
            break
    return lhat, stop, table


def np_metrics_at_lambda(outputs, labels, lam, seed=None):
    """calibrate_model.py:31-60 (get_rcps_metrics_from_outputs) restated on numpy fp32 + the reference's own RNG calls.

    Batches of 64 in order (:35), per batch one ``np.random.choice(px, size=b)`` (:44); then one ``torch.rand(N)`` (:51).
    ``seed`` (optional) seeds both generators the way tests/golden/make_golden_metrics.py does.
    Returns (losses (N,) fp32, sizes (N,) fp32 tensor, spearman, stratified_risks (4,) tensor, mse, spatial (H,W))."""
    import torch
    from scipy.stats import spearmanr
    outputs = np.asarray(outputs, dtype=np.float32)
    labels = np.asarray(labels, dtype=np.float32)
    if seed is not None:
        np.random.seed(seed)
        torch.manual_seed(seed)
    n = outputs.shape[0]
    losses, sizes, residuals, maps = [], [], [], []
    # :35 iterates a DataLoader: its iterator draws one int64 "base seed" from torch's default generator before the first
    # batch (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__), which shifts the torch.rand of :50
    torch.empty((), dtype=torch.int64).random_()
    for lo in range(0, n, 64):
        x, y = outputs[lo:lo + 64], labels[lo:lo + 64]
        lower, pred, upper = np_nested_sets(x, lam)                                   # :40 (lam = model.lhat)
        losses.append(np_fraction_missed((lower, pred, upper), y)[0])                 # :41
        b = x.shape[0]
        full = (upper - lower).reshape(b, -1)                                         # :42
        idx = np.random.choice(full.shape[1], size=b)                                 # :43
        sizes.append(full[np.arange(b), idx])                                         # :44
        with np.errstate(all="ignore"):
            residuals.append(np.abs(y - pred).reshape(b, -1)[np.arange(b), idx])      # :45
            maps.append((y > upper).astype(np.float32) + (y < lower).astype(np.float32))   # :46
    losses = torch.from_numpy(np.concatenate(losses))
    sizes = torch.from_numpy(np.concatenate(sizes))
    sizes = sizes + torch.rand(size=sizes.shape) * 1e-6                               # :50
    residuals = np.concatenate(residuals)
    spearman = spearmanr(residuals, sizes)[0]                                         # :52
    mse = (residuals * residuals).mean().item()                                       # :53
    spatial = np.concatenate(maps, axis=0).mean(axis=0).mean(axis=0)                  # :54
    size_bins = torch.tensor([0, torch.quantile(sizes, 0.25), torch.quantile(sizes, 0.5), torch.quantile(sizes, 0.75)])
    buckets = torch.bucketize(sizes, size_bins) - 1                                   # :56
    strat = torch.tensor([losses[buckets == bucket].mean() for bucket in range(size_bins.shape[0])])
    return losses, sizes, spearman, strat, mse, spatial
