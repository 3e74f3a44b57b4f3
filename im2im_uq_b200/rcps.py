"""Tensor-level wrappers over the RCPS entry points of the C ABI (include/im2im_uq.h).

All tensors are CUDA fp32; outputs follow the reference's head layouts and labels are (N, C, H, W):
  head QUANTILES / SOFTMAX_SETS   (N, 3, C, H, W) = (lower, pred, upper)      quantile_layer.py:19-21
  head RESIDUAL / GAUSSIAN        (N, 2, C, H, W) = (pred, width | variance)  residual_magnitude_layer.py:17-19,
                                                                              gaussian_layer.py:17-19
"""
from typing import Optional, Tuple

import torch

from . import _lib


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.Im2ImError(f"{name} must be a CUDA tensor: im2im_uq_b200 has no CPU path (got device {t.device})")
    if t.dtype != torch.float32:
        raise _lib.Im2ImError(f"{name} must be float32 (got {t.dtype})")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _score_planes(outputs: torch.Tensor, labels: Optional[torch.Tensor], head: int = _lib.IM2IM_HEAD_QUANTILES):
    """Raw pointers/strides of the (lower, pred, upper[, label]) planes without copying when the layout allows.
    For the two-plane heads the `lower` slot repeats the prediction plane (the ABI ignores it)."""
    _require_cuda(outputs, "outputs")
    three = head in _lib.THREE_PLANE_HEADS
    if outputs.dim() < 3 or outputs.shape[1] != (3 if three else 2):
        want = "(N, 3, ...) = (lower, pred, upper)" if three else "(N, 2, ...) = (pred, width)"
        raise _lib.Im2ImError(f"outputs must be {want} for head kind {head}; got {tuple(outputs.shape)}")
    n = outputs.shape[0]
    px = 1
    for s in outputs.shape[2:]:
        px *= s
    inner_contig = outputs[0, 0].is_contiguous() if n > 0 and px > 0 else True
    if not inner_contig:
        outputs = outputs.contiguous()
    s_img, s_plane = (outputs.stride(0), outputs.stride(1)) if n > 0 and px > 0 else (outputs.shape[1] * px, px)
    base = outputs.data_ptr()
    ptrs = [base, base + 4 * s_plane, base + 8 * s_plane] if three else [base, base, base + 4 * s_plane]
    strides = [s_img, s_img, s_img]
    keep = [outputs]
    if labels is not None:
        _require_cuda(labels, "labels")
        if labels.shape[0] != n or labels.numel() != n * px:
            raise _lib.Im2ImError(f"labels {tuple(labels.shape)} do not match outputs {tuple(outputs.shape)}")
        if n > 0 and px > 0 and not labels[0].is_contiguous():
            labels = labels.contiguous()
        ptrs.append(labels.data_ptr())
        strides.append(labels.stride(0) if n > 0 and px > 0 else px)
        keep.append(labels)
    return n, px, ptrs, strides, keep


def miss_counts(outputs: torch.Tensor, labels: torch.Tensor, lambdas_sorted: torch.Tensor,
                counts: Optional[torch.Tensor] = None, totals: Optional[torch.Tensor] = None,
                zero: bool = True, force_generic: bool = False,
                head: int = _lib.IM2IM_HEAD_QUANTILES) -> Tuple[torch.Tensor, torch.Tensor]:
    """counts[i, j] = #pixels of image i missed at lambdas_sorted[j]; totals[j] += column sums (int64).

    One pass over HBM for the whole grid (im2im_rcps_miss_counts).  ``lambdas_sorted`` is a CUDA fp32 vector,
    finite and ascending.  ``counts``/``totals`` may be preallocated (e.g. to accumulate chunks with zero=False).
    """
    lib = _lib.load()
    n, px, ptrs, strides, keep = _score_planes(outputs, labels, head)
    _require_cuda(lambdas_sorted, "lambdas_sorted")
    lambdas_sorted = lambdas_sorted.contiguous()
    n_lam = lambdas_sorted.numel()
    dev = outputs.device
    if counts is None:
        counts = torch.empty((n, n_lam), dtype=torch.int32, device=dev)
        zero = True
    if totals is None:
        totals = torch.empty((n_lam,), dtype=torch.int64, device=dev)
        zero = True
    assert counts.is_contiguous() and counts.dtype == torch.int32 and tuple(counts.shape) == (n, n_lam)
    assert totals.is_contiguous() and totals.dtype == torch.int64 and totals.numel() == n_lam
    flags = (_lib.IM2IM_RCPS_ZERO_OUTPUTS if zero else 0) | (_lib.IM2IM_RCPS_FORCE_GENERIC if force_generic else 0)
    with torch.cuda.device(dev):
        rc = lib.im2im_rcps_miss_counts(ptrs[0], ptrs[1], ptrs[2], ptrs[3], n, px, strides[0], strides[1], strides[2],
                                        strides[3], lambdas_sorted.data_ptr(), n_lam, head,
                                        counts.data_ptr(), totals.data_ptr(), flags, _stream_ptr(dev))
    _lib.check(rc, "im2im_rcps_miss_counts")
    del keep
    return counts, totals


def counts_from_hist(hist: torch.Tensor, counts: torch.Tensor, totals: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-image rank histograms (uint32/int32 [n, L+1], written by the head-fused convolution epilogue,
    conv.head_conv_tc_hist) -> counts int32 [n, L] as ``miss_counts`` writes them; ``totals`` (int64 [L]) += column sums.
    The histogram is zero again afterwards (im2im_rcps_counts_from_hist)."""
    lib = _lib.load()
    if not hist.is_cuda:
        raise _lib.Im2ImError(f"hist must be a CUDA tensor: im2im_uq_b200 has no CPU path (got device {hist.device})")
    n, l1 = hist.shape
    assert hist.is_contiguous() and hist.dtype in (torch.int32, torch.uint32)
    assert counts.is_cuda and counts.is_contiguous() and counts.dtype == torch.int32 and tuple(counts.shape) == (n, l1 - 1)
    if totals is not None:
        assert totals.is_cuda and totals.is_contiguous() and totals.dtype == torch.int64 and totals.numel() == l1 - 1
    with torch.cuda.device(hist.device):
        rc = lib.im2im_rcps_counts_from_hist(hist.data_ptr(), n, l1 - 1, counts.data_ptr(),
                                             totals.data_ptr() if totals is not None else None, _stream_ptr(hist.device))
    _lib.check(rc, "im2im_rcps_counts_from_hist")
    return counts


def loss_table(counts: torch.Tensor, px: int, first_visited_col: int = 0, out: Optional[torch.Tensor] = None,
               first_visited_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 table[i,j] = float(counts[i,j])/float(px); columns below first_visited_col are zero.

    ``first_visited_dev`` (int32 CUDA tensor, first element used) takes the column from device memory instead, e.g.
    ``result[3:]`` of im2im_rcps_decide, so no host round trip is needed between the decision and the table."""
    lib = _lib.load()
    assert counts.is_cuda and counts.dtype == torch.int32 and counts.is_contiguous() and counts.dim() == 2
    n, n_lam = counts.shape
    if out is None:
        out = torch.empty((n, n_lam), dtype=torch.float32, device=counts.device)
    with torch.cuda.device(counts.device):
        if first_visited_dev is not None:
            assert first_visited_dev.is_cuda and first_visited_dev.dtype == torch.int32
            rc = lib.im2im_rcps_loss_table_dev(counts.data_ptr(), n, n_lam, px, first_visited_dev.data_ptr(),
                                               out.data_ptr(), _stream_ptr(counts.device))
        else:
            rc = lib.im2im_rcps_loss_table(counts.data_ptr(), n, n_lam, px, first_visited_col, out.data_ptr(),
                                           _stream_ptr(counts.device))
    _lib.check(rc, "im2im_rcps_loss_table")
    return out


def quantile_nested_sets(outputs: torch.Tensor, lam: float, write_back_clamp: bool = True):
    """(lower_edge, prediction, upper_edge) at one lambda; prediction is a view of outputs[:, 1].

    write_back_clamp reproduces the reference's in-place clamp of ``outputs`` (quantile_layer.py:39-40).
    """
    lib = _lib.load()
    n, px, ptrs, strides, keep = _score_planes(outputs, None)
    src = keep[0]
    if write_back_clamp and src.data_ptr() != outputs.data_ptr():
        write_back_clamp = False  # a contiguous copy was made; nothing the caller could observe
    shape = tuple(outputs.shape[:1]) + tuple(outputs.shape[2:])
    lower = torch.empty(shape, dtype=torch.float32, device=outputs.device)
    upper = torch.empty(shape, dtype=torch.float32, device=outputs.device)
    with torch.cuda.device(outputs.device):
        rc = lib.im2im_quantile_nested_sets(ptrs[0], ptrs[1], ptrs[2], n, px, strides[0], strides[1], strides[2],
                                            float(lam), 1 if write_back_clamp else 0, lower.data_ptr(),
                                            upper.data_ptr(), _stream_ptr(outputs.device))
    _lib.check(rc, "im2im_quantile_nested_sets")
    return lower, src[:, 1], upper


def head_nested_sets(outputs: torch.Tensor, lam: float, head: int):
    """(lower_edge, prediction, upper_edge) at one lambda for any head kind (im2im_nested_sets): the head's set
    function followed by the +/-1e-6 clamp of add_uncertainty.py:35-36.  prediction is a view of ``outputs``."""
    if head == _lib.IM2IM_HEAD_QUANTILES:
        return quantile_nested_sets(outputs, lam, write_back_clamp=False)
    lib = _lib.load()
    n, px, ptrs, strides, keep = _score_planes(outputs, None, head)
    src = keep[0]
    shape = tuple(outputs.shape[:1]) + tuple(outputs.shape[2:])
    lower = torch.empty(shape, dtype=torch.float32, device=outputs.device)
    upper = torch.empty(shape, dtype=torch.float32, device=outputs.device)
    with torch.cuda.device(outputs.device):
        rc = lib.im2im_nested_sets(head, ptrs[0], ptrs[1], ptrs[2], n, px, strides[0], strides[1], strides[2],
                                   float(lam), 0, lower.data_ptr(), upper.data_ptr(), _stream_ptr(outputs.device))
    _lib.check(rc, "im2im_nested_sets")
    return lower, src[:, 1 if head in _lib.THREE_PLANE_HEADS else 0], upper


def softmax_sets(logits: torch.Tensor) -> torch.Tensor:
    """Softmax head logits (N, K, ...) -> (N, 3, ...) = (lower quantile, argmax prediction, upper quantile): the
    lambda-independent half of softmax_nested_sets_from_output (softmax_layer.py:34-48), computed once."""
    lib = _lib.load()
    _require_cuda(logits, "logits")
    if logits.dim() < 3:
        raise _lib.Im2ImError(f"logits must be (N, K, ...); got {tuple(logits.shape)}")
    logits = logits.contiguous()
    n, k = logits.shape[:2]
    inner = 1
    for s_ in logits.shape[2:]:
        inner *= s_
    sets = torch.empty((n, 3) + tuple(logits.shape[2:]), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = lib.im2im_softmax_sets(logits.data_ptr(), n, k, inner, k * inner, inner, sets.data_ptr(),
                                    _stream_ptr(logits.device))
    _lib.check(rc, "im2im_softmax_sets")
    return sets


def miss_map(outputs: torch.Tensor, labels: torch.Tensor, lam: float,
             head: int = _lib.IM2IM_HEAD_QUANTILES) -> torch.Tensor:
    """int32 map over (C,H,W): number of images whose pixel is missed at ``lam``."""
    lib = _lib.load()
    n, px, ptrs, strides, keep = _score_planes(outputs, labels, head)
    out = torch.empty(tuple(outputs.shape[2:]), dtype=torch.int32, device=outputs.device)
    with torch.cuda.device(outputs.device):
        rc = lib.im2im_rcps_miss_map(ptrs[0], ptrs[1], ptrs[2], ptrs[3], n, px, strides[0], strides[1], strides[2],
                                     strides[3], float(lam), head, out.data_ptr(),
                                     _lib.IM2IM_RCPS_ZERO_OUTPUTS, _stream_ptr(outputs.device))
    _lib.check(rc, "im2im_rcps_miss_map")
    del keep
    return out


def fraction_missed_counts(lower_edge: torch.Tensor, upper_edge: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """int32 miss count per image for already computed endpoints (any (B, ...) shapes with equal numel per image)."""
    lib = _lib.load()
    for name, t in (("lower_edge", lower_edge), ("upper_edge", upper_edge), ("label", label)):
        _require_cuda(t, name)
    n = label.shape[0]
    lower_edge = lower_edge.reshape(n, -1)
    upper_edge = upper_edge.reshape(n, -1)
    label = label.reshape(n, -1)
    px = label.shape[1]
    if lower_edge.shape != label.shape or upper_edge.shape != label.shape:
        raise _lib.Im2ImError("fraction_missed_counts: endpoint and label shapes differ")
    if px > 0 and n > 0:
        lower_edge = lower_edge if lower_edge.stride(1) == 1 else lower_edge.contiguous()
        upper_edge = upper_edge if upper_edge.stride(1) == 1 else upper_edge.contiguous()
        label = label if label.stride(1) == 1 else label.contiguous()
    counts = torch.empty((n,), dtype=torch.int32, device=label.device)
    with torch.cuda.device(label.device):
        for lo in range(0, max(n, 1), 65535):
            hi = min(n, lo + 65535)
            if hi <= lo:
                break
            rc = lib.im2im_fraction_missed_counts(lower_edge[lo:hi].data_ptr(), upper_edge[lo:hi].data_ptr(),
                                                  label[lo:hi].data_ptr(), hi - lo, px, lower_edge.stride(0),
                                                  upper_edge.stride(0), label.stride(0), counts[lo:hi].data_ptr(),
                                                  _stream_ptr(label.device))
            _lib.check(rc, "im2im_fraction_missed_counts")
    return counts
