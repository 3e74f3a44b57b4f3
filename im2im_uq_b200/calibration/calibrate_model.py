"""RCPS calibration with the reference's entry points, running on hand-written sm_100a kernels.

Mirror of ``core/calibration/calibrate_model.py`` of aangelopoulos/im2im-uq - same function names, argument meaning,
return types and error behaviour:

    get_rcps_losses_from_outputs(model, out_dataset, rcps_loss_fn, lam, device)      reference :21-29
    get_rcps_metrics_from_outputs(model, out_dataset, rcps_loss_fn, device)          reference :31-60
    evaluate_from_loss_table(loss_table, n, alpha, delta)                            reference :62-74
    fraction_missed_loss(pset, label)                                                reference :76-80
    get_rcps_loss_fn(config)                                                         reference :82-87
    calibrate_model(model, dataset, config) -> (model, calib_loss_table)             reference :89-145

What is different underneath: the scores stay resident in HBM, every lambda column is produced by ONE pass of
``im2im_rcps_miss_counts`` (include/im2im_uq.h), and the stopping rule is replayed on the host from exact integer
counts (sweep.py).  There is no CPU path: ``config['device']`` must name a CUDA device.

Deliberate deviation: the reference's ``fraction_missed_loss`` squeezes away a batch of one (N % 64 == 1 makes
``calibrate_model`` raise on a shape mismatch); here every image always yields one loss.
"""
from typing import Optional, Tuple

import numpy as np
import torch
from scipy.stats import spearmanr
from torch.utils.data import DataLoader, TensorDataset

from .. import _lib, rcps
from . import sweep
from .bounds import HB_mu_plus, hb_stop_bracket

_CHUNK_BYTES = 2 << 30  # host->device staging granularity for CPU-resident score tensors


def _cuda_device(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.Im2ImError(f"im2im_uq_b200 runs its hot path on CUDA only (config['device']={device!r}); "
                              "there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


# ------------------------------------------------------------------------------------------------ loss functions
def fraction_missed_loss(pset, label):
    """Per-image fraction of pixels outside [lower, upper]: ``(lower>y)+(upper<y)``, clipped to 1, mean over pixels.

    pset = (lower_edge, prediction, upper_edge) CUDA tensors of shape (B, ...); returns a (B,) fp32 CUDA tensor equal
    to float(count)/float(pixels) - exactly what the reference's fp32 mean of 0/1 values produces.
    """
    counts = rcps.fraction_missed_counts(pset[0], pset[2], label)
    px = label[0].numel() if label.shape[0] > 0 else 1
    return rcps.loss_table(counts.reshape(-1, 1), max(px, 1))[:, 0]


def get_rcps_loss_fn(config):
    string = config['rcps_loss']
    if string == 'fraction_missed':
        return fraction_missed_loss
    else:
        raise NotImplementedError


# ------------------------------------------------------------------------------------------------ score access
def _dataset_tensors(out_dataset) -> Tuple[torch.Tensor, torch.Tensor]:
    if isinstance(out_dataset, TensorDataset) and len(out_dataset.tensors) == 2:
        return out_dataset.tensors
    if isinstance(out_dataset, (tuple, list)) and len(out_dataset) == 2 and torch.is_tensor(out_dataset[0]):
        return out_dataset[0], out_dataset[1]
    # generic map-style dataset of (output, label) pairs
    loader = DataLoader(out_dataset, batch_size=64, shuffle=False, num_workers=0)
    xs, ys = zip(*[(b[0], b[1]) for b in loader])
    return torch.cat(xs, dim=0), torch.cat(ys, dim=0)


def _fused_head(model, rcps_loss_fn=None):
    """(head kind, scores_from_output or None) when the model's set function is one of the built-in heads (each tags
    itself with the head kind its kernel uses) and the loss is fraction_missed; None for user-supplied functions."""
    if rcps_loss_fn is not None and rcps_loss_fn is not fraction_missed_loss:
        return None
    fn = getattr(model, "in_nested_sets_from_output_fn", None)
    kind = getattr(fn, "im2im_head_kind", None)
    if kind is None:
        return None
    return kind, getattr(fn, "im2im_scores_from_output", None)


def _head_scores(model, outputs, device):
    """Head kind + the score planes the sweep kernels read.  For the softmax head the lambda-independent half of the set
    function (softmax, cumsum, quantiles; softmax_layer.py:34-48) is evaluated ONCE here, in chunks, on the device."""
    fused = _fused_head(model)
    if fused is None:
        raise NotImplementedError("calibration needs one of the built-in heads (add_uncertainty) - a user-supplied "
                                  "nested-set function has no one-pass kernel")
    kind, to_scores = fused
    if to_scores is None:
        return kind, outputs
    per_image = max(outputs[0].numel() * 4, 1) if outputs.shape[0] else 1
    parts = [to_scores(outputs[lo:hi].to(device)) for lo, hi in _chunks(outputs.shape[0], per_image)]
    if not parts:
        return kind, torch.empty((0, 3) + tuple(outputs.shape[2:]), dtype=torch.float32, device=device)
    return kind, parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)


def _resolve_lambda(model, lam):
    if lam is None:
        if model.lhat is None:
            raise Exception("You have to specify lambda unless your model is already calibrated.")
        lam = model.lhat
    return lam


def _chunks(n: int, bytes_per_image: int):
    step = max(1, _CHUNK_BYTES // max(bytes_per_image, 1))
    for lo in range(0, n, step):
        yield lo, min(n, lo + step)


def _miss_counts_any_device(outputs, labels, lam_sorted_dev, device, counts=None, totals=None,
                            head=_lib.IM2IM_HEAD_QUANTILES):
    """One-pass miss counts for scores that live on `device` already, or on the host (staged in pinned chunks)."""
    n = outputs.shape[0]
    n_lam = lam_sorted_dev.numel()
    if counts is None:
        counts = torch.zeros((n, n_lam), dtype=torch.int32, device=device)
    if totals is None:
        totals = torch.zeros((n_lam,), dtype=torch.int64, device=device)
    if outputs.is_cuda and labels.is_cuda:
        rcps.miss_counts(outputs, labels, lam_sorted_dev, counts=counts, totals=totals, zero=False, head=head)
        return counts, totals
    per_image = (outputs[0].numel() + labels[0].numel()) * 4 if n else 1
    copy_stream = _copy_stream(device)
    main = torch.cuda.current_stream(device)
    prev = None
    for lo, hi in _chunks(n, per_image):
        with torch.cuda.stream(copy_stream):
            x = outputs[lo:hi].to(device, non_blocking=True)
            y = labels[lo:hi].to(device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        main.wait_event(ready)
        rcps.miss_counts(x, y, lam_sorted_dev, counts=counts[lo:hi], totals=totals, zero=False, head=head)
        x.record_stream(main)
        y.record_stream(main)
        prev = (x, y)
    del prev
    return counts, totals


# ------------------------------------------------------------------------------------------------ reference API
def get_rcps_losses_from_outputs(model, out_dataset, rcps_loss_fn, lam, device):
    """Per-image RCPS losses at ONE lambda; returns an (N,) fp32 CPU tensor like the reference."""
    device = _cuda_device(device)
    model = model.to(device)
    outputs, labels = _dataset_tensors(out_dataset)
    lam = _resolve_lambda(model, lam)
    with torch.no_grad():
        if _fused_head(model, rcps_loss_fn) is not None:
            kind, scores = _head_scores(model, outputs, device)
            lam_dev = torch.as_tensor(lam, dtype=torch.float32).reshape(1).to(device)
            counts, _ = _miss_counts_any_device(scores, labels, lam_dev, device, head=kind)
            px = labels[0].numel() if labels.shape[0] else 1
            return rcps.loss_table(counts, max(px, 1))[:, 0].cpu()
        losses = []
        for lo in range(0, outputs.shape[0], 64):
            x = outputs[lo:lo + 64].to(device).clone()
            sets = model.nested_sets_from_output(x, lam)
            losses = losses + [rcps_loss_fn(sets, labels[lo:lo + 64].to(device)).cpu(), ]
        return torch.cat(losses, dim=0)


def get_rcps_metrics_from_outputs(model, out_dataset, rcps_loss_fn, device):
    """Risk, sampled set sizes, Spearman(size, residual), size-stratified risk, mse, (H,W) spatial miscoverage at lhat.

    RNG parity: the reference draws ``np.random.choice(pixels, size=batch)`` once per batch of 64, in order, then one
    ``torch.rand(N)`` for the jitter (:44,:51); the same calls are made here in the same order.
    """
    device = _cuda_device(device)
    model = model.to(device)
    outputs, labels = _dataset_tensors(out_dataset)
    lam = _resolve_lambda(model, None)
    n = outputs.shape[0]
    px = labels[0].numel()
    with torch.no_grad():
        if _fused_head(model, rcps_loss_fn) is None:
            raise NotImplementedError("metrics are implemented for the built-in heads + fraction_missed loss")
        kind, outputs_d = _head_scores(model, outputs, device)
        planes = outputs_d.shape[1]
        outputs_d = outputs_d if outputs_d.is_cuda else outputs_d.to(device)
        labels_d = labels if labels.is_cuda else labels.to(device)
        lam_dev = torch.as_tensor(lam, dtype=torch.float32).reshape(1).to(device)
        counts, _ = rcps.miss_counts(outputs_d, labels_d, lam_dev, head=kind)
        losses = rcps.loss_table(counts, px)[:, 0]
        # the reference iterates a DataLoader (:35-36); its iterator draws one int64 base seed from torch's default
        # generator (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__), which shifts the torch.rand below
        torch.empty((), dtype=torch.int64).random_()
        # one random pixel per image, drawn batch by batch like the reference
        idx_parts = []
        for lo in range(0, n, 64):
            b = min(64, n - lo)
            idx_parts.append(np.random.choice(px, size=b))
        idx = torch.from_numpy(np.concatenate(idx_parts)).to(device)
        rows = torch.arange(n, device=device)
        picked = outputs_d.reshape(n, planes, px)[rows, :, idx].reshape(n, planes, 1).contiguous()
        lo_e, pred_e, up_e = rcps.head_nested_sets(picked, float(lam), kind)
        sizes = (up_e - lo_e).reshape(n).cpu()
        residuals = (labels_d.reshape(n, px)[rows, idx] - pred_e.reshape(n)).abs()
        miss_map = rcps.miss_map(outputs_d, labels_d, float(lam), head=kind)
    sizes = sizes + torch.rand(size=sizes.shape).to(sizes.device) * 1e-6
    residuals = residuals.detach().cpu().numpy()
    spearman = spearmanr(residuals, sizes)[0]
    mse = (residuals * residuals).mean().item()
    # numpy fp32 mean over images (exact integer sum / N), then over the channel axis - as the reference does on host
    spatial_miscoverage = (miss_map.cpu().numpy().astype(np.float32) / np.float32(n)).reshape(labels.shape[1:]).mean(axis=0)
    size_bins = torch.tensor([0, torch.quantile(sizes, 0.25), torch.quantile(sizes, 0.5), torch.quantile(sizes, 0.75)])
    buckets = torch.bucketize(sizes, size_bins) - 1
    losses_c = losses.cpu()
    stratified_risks = torch.tensor([losses_c[buckets == bucket].mean() for bucket in range(size_bins.shape[0])])
    return losses, sizes, spearman, stratified_risks, mse, spatial_miscoverage


def evaluate_from_loss_table(loss_table, n, alpha, delta):
    """Random calibration/validation split of a saved loss table; returns the validation risk at the first lambda
    whose HB bound is <= delta (the reference compares against ``delta`` here, :70 - kept as is).

    Host-only like the reference.  The per-column bound is screened through the cached level set of HB_mu_plus and
    evaluated exactly only around the decision, instead of one brentq solve per column.
    """
    with torch.no_grad():
        perm = torch.randperm(loss_table.shape[0])
        loss_table = loss_table[perm]
        calib_table, val_table = loss_table[:n], loss_table[n:]
        Rhats = calib_table.mean(dim=0)
        idx_lambda = _first_accepted_column(Rhats, n, delta)
        if idx_lambda is None:
            print("No rejections made!")
            idx_lambda = 0
        else:
            # the reference indexes with the 1-element tensor `nonzero()[0]` (:70,:74): an index_select copy of shape
            # (N_val, 1), whose fp32 mean is summed in a different order than the strided view an int index gives
            idx_lambda = torch.tensor([idx_lambda])
        return val_table[:, idx_lambda].mean()


def _first_accepted_column(Rhats: torch.Tensor, n: int, delta: float) -> Optional[int]:
    """Smallest j with HB_mu_plus(Rhats[j], n, delta) <= delta, or None."""
    r = Rhats.double().numpy()
    r_lo, r_hi = hb_stop_bracket(int(n), float(delta), float(delta))  # level set of HB > delta
    slack = 1e-9
    for j in range(r.shape[0]):
        m = float(Rhats[j])
        if m > 0 and np.isfinite(r_hi) and m > r_hi + slack:
            continue  # certainly HB > delta
        if m > 0 and np.isfinite(r_lo) and m < r_lo - slack:
            return j  # certainly HB <= delta
        if HB_mu_plus(Rhats[j], n, delta) <= delta:
            return j
    return None


# ------------------------------------------------------------------------------------------------ the sweep
def device_decide_fn(px: int, config: dict, result: Optional[torch.Tensor] = None):
    """f(totals, n_total) -> (stop, decided): the stopping-rule screen on the device (im2im_rcps_decide) followed by a
    16-byte device->host read.  ``result`` (int32[4], CUDA) is left holding the kernel's output, so a caller can feed
    result[3:] to ``rcps.loss_table(first_visited_dev=...)`` before the read."""
    lib = _lib.load()

    def decide(totals: torch.Tensor, n_total: int, read: bool = True):
        nonlocal result
        if result is None:
            result = torch.empty(4, dtype=torch.int32, device=totals.device)
        n_px, gamma, alpha32, r_lo, r_hi, slack = sweep.screening_constants(n_total, px, config['alpha'], config['delta'])
        with torch.cuda.device(totals.device):
            rc = lib.im2im_rcps_decide(totals.data_ptr(), totals.numel(), n_px, gamma, alpha32, r_lo, r_hi, slack,
                                       result.data_ptr(), torch.cuda.current_stream(totals.device).cuda_stream)
        _lib.check(rc, "im2im_rcps_decide")
        if not read:
            return result
        host = result.cpu()
        return int(host[0]), bool(host[1])

    return decide


def _peer_timeout_s() -> float:
    """Bound on the in-kernel wait for the peers' flags (seconds; 0 = wait for ever).  Two minutes by default - the ranks
    call ``RcpsGraph.run()`` in lockstep, so this only has to absorb host-side skew (a slow loader, a first-call JIT), and
    a spinning kernel must not outlive a crashed peer for long; raise it (or set 0) with IM2IM_P2P_TIMEOUT_S when a
    debugger or a very uneven pipeline sits between the ranks."""
    import os
    return float(os.environ.get("IM2IM_P2P_TIMEOUT_S", "120"))


def peer_memory_available(group) -> bool:
    """Cheap LOCAL capability check made before any collective allocation / rendezvous: torch symmetric memory importable
    and every device of the group peer-accessible.  Ranks agree on the outcome with one all-reduce (RcpsGraph)."""
    try:
        import torch.distributed._symmetric_memory as symm  # noqa: F401
        dev = torch.cuda.current_device()
        return all(d == dev or torch.cuda.can_device_access_peer(dev, d) for d in range(torch.cuda.device_count()))
    except Exception:  # noqa: BLE001
        return False


class PeerTotals:
    """Peer-mapped buffers for the totals exchange: every rank's mailbox uint64[2][world][L] and flag array uint32[world]
    in torch symmetric memory (CUDA peer / fabric mappings over NVLink), plus this rank's device-side epoch (used by
    ``im2im_rcps_decide_p2p``; the fused kernel keeps its epoch in its own workspace).  Construction is collective over
    ``group``: call it only after the ranks have AGREED that peer memory is available (``peer_memory_available`` + an
    all-reduce) - an exception on one rank in here would leave the others waiting in the rendezvous."""

    def __init__(self, n_lambdas: int, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.mailbox = symm.empty(2 * self.world * n_lambdas, dtype=torch.int64, device=device)
        self.flags = symm.empty(max(self.world, 32), dtype=torch.int32, device=device)
        self.mailbox.zero_()
        self.flags.zero_()
        name = group.group_name if hasattr(group, "group_name") else group
        h_mail = symm.rendezvous(self.mailbox, name)
        h_flag = symm.rendezvous(self.flags, name)
        self.mail_ptrs = torch.tensor([int(p) for p in h_mail.buffer_ptrs], dtype=torch.int64, device=device)
        self.flag_ptrs = torch.tensor([int(p) for p in h_flag.buffer_ptrs], dtype=torch.int64, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self._handles = (h_mail, h_flag)
        torch.cuda.synchronize(device)
        dist.barrier(group)                 # every rank's flags are zero before anyone publishes

    def close(self):
        self._handles = None
        self.mailbox = self.flags = None


class RcpsGraph:
    """The device side of one calibration, captured ONCE into a CUDA graph and replayed.

    Fused path (default whenever the bulk-copy kernel applies): ONE kernel per calibration
    (``im2im_rcps_calibrate_fused``): miss counts written without a memset, block totals reduced in a self-cleaning
    workspace, the last block all-reduces them over NVLink peer memory (multi-GPU), screens the stopping rule, publishes the
    16-byte result into mapped pinned host memory, and the loss table is written by the same launch.  The host spins on
    the result's epoch tag instead of synchronising the stream.
    Fallback path: memsets + ``im2im_rcps_miss_counts`` + (NCCL all-reduce | ``im2im_rcps_decide_p2p``) +
    ``im2im_rcps_loss_table_dev`` + a 16-byte copy.

    Scores must stay resident and unchanged in shape; their contents may change between replays.

        plan = RcpsGraph(outputs, labels, config, group=None, n_total=None)
        lhat, stop, decided = plan.run()        # decided False -> a column fell in the guard band, call plan.replay_on_host()
        plan.close()                            # before destroying the process group
    """

    def __init__(self, outputs: torch.Tensor, labels: torch.Tensor, config: dict, group=None, n_total=None,
                 head: int = _lib.IM2IM_HEAD_QUANTILES, p2p: bool = True, fused: bool = True):
        assert outputs.is_cuda and labels.is_cuda
        self.config, self.group, self.head = config, group, head
        self.outputs, self.labels = outputs, labels
        dev = outputs.device
        lib = _lib.load()
        self.lambdas, self.dlambda, lam_prime, self.default_lhat = sweep.lambda_grid(config)
        if not bool((lam_prime[1:] >= lam_prime[:-1]).all()) or not bool(torch.isfinite(lam_prime).all()):
            raise ValueError("RcpsGraph needs a finite ascending lambda grid")
        self.lam_dev = lam_prime.to(dev)
        n, L = outputs.shape[0], lam_prime.numel()
        self.px = labels[0].numel()
        self.n_total = n_total if n_total is not None else n
        self.counts = torch.empty((n, L), dtype=torch.int32, device=dev)
        self.totals = torch.empty((L,), dtype=torch.int64, device=dev)
        self.table = torch.empty((n, L), dtype=torch.float32, device=dev)
        self.result = torch.empty(4, dtype=torch.int32, device=dev)
        self.result_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._result_np = self.result_host.numpy()      # same memory; numpy scalar reads cost ~0.1 us (tensor indexing ~3 us)
        self._decide = device_decide_fn(self.px, config, result=self.result)
        self.peer = None
        self.graph = None
        self._launches = 0          # executions of the fused kernel so far = the epoch tag the host waits for
        self._planes = rcps._score_planes(outputs, labels, head)
        n_, px_, ptrs, strides, _keep = self._planes
        self.fused = bool(fused) and n > 0 and lib.im2im_rcps_calibrate_fused_check(
            ptrs[0], ptrs[1], ptrs[2], ptrs[3], n_, px_, strides[0], strides[1], strides[2], strides[3], L, head) == 0
        world = 1
        if group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(group)
            # all ranks must take the same path: agree BEFORE any collective allocation or rendezvous
            ok = torch.tensor([1 if (p2p and peer_memory_available(group)) else 0, 1 if self.fused else 0],
                              dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            use_peer, self.fused = bool(ok[0]), bool(ok[1])
            if use_peer:
                self.peer = PeerTotals(L, group, dev)      # collective; a failure in here is fatal for the job
                self.local_totals = torch.empty((L,), dtype=torch.int64, device=dev)
            else:
                self.fused = False                         # the fused kernel exchanges over peer memory only
        self.world = world
        if self.fused:
            ws_bytes = int(lib.im2im_rcps_fused_workspace_bytes(L))
            self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)   # zero-filled ONCE
        try:
            self._enqueue()                  # warm-up outside capture (lazy module loads, NCCL channel setup)
            torch.cuda.synchronize(dev)
        except _lib.Im2ImError:
            if not (self.fused and world == 1):
                raise
            # e.g. the cooperative launch was refused (not every block can be resident: MPS / a shared GPU): the separate
            # kernels need no co-residency
            self.fused = False
            self._launches = 0
            self._enqueue()
            torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self._enqueue(count=False)
        self.kernels_per_replay = _lib.launch_count() - before  # libim2im_uq kernels inside one replay

    def _enqueue(self, count: bool = True):
        dev = self.totals.device
        n_px, gamma, alpha32, r_lo, r_hi, slack = sweep.screening_constants(self.n_total, self.px,
                                                                            self.config['alpha'], self.config['delta'])
        if self.fused:
            n_, px_, ptrs, strides, _keep = self._planes
            peer = self.peer
            with torch.cuda.device(dev):
                rc = _lib.load().im2im_rcps_calibrate_fused(
                    ptrs[0], ptrs[1], ptrs[2], ptrs[3], n_, px_, strides[0], strides[1], strides[2], strides[3],
                    self.lam_dev.data_ptr(), self.lam_dev.numel(), self.head, self.counts.data_ptr(),
                    self.table.data_ptr(), self.totals.data_ptr(), n_px, gamma, alpha32, r_lo, r_hi, slack,
                    self.workspace.data_ptr(), self.workspace.numel(),
                    peer.mail_ptrs.data_ptr() if peer is not None else None,
                    peer.flag_ptrs.data_ptr() if peer is not None else None,
                    peer.rank if peer is not None else 0, peer.world if peer is not None else 1, _peer_timeout_s(),
                    self.result.data_ptr(), self.result_host.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "im2im_rcps_calibrate_fused")
            if count:
                self._launches += 1
            return
        if self.peer is not None:
            # totals stay local; the reduction over ranks happens inside the decision kernel, over peer memory
            rcps.miss_counts(self.outputs, self.labels, self.lam_dev, counts=self.counts, totals=self.local_totals,
                             zero=True, head=self.head)
            with torch.cuda.device(dev):
                rc = _lib.load().im2im_rcps_decide_p2p(
                    self.local_totals.data_ptr(), self.peer.mail_ptrs.data_ptr(), self.peer.flag_ptrs.data_ptr(),
                    self.peer.epoch.data_ptr(), self.peer.rank, self.peer.world, self.totals.numel(), n_px, gamma, alpha32,
                    r_lo, r_hi, slack, self.totals.data_ptr(), self.result.data_ptr(),
                    torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "im2im_rcps_decide_p2p")
        else:
            rcps.miss_counts(self.outputs, self.labels, self.lam_dev, counts=self.counts, totals=self.totals, zero=True,
                             head=self.head)
            if self.group is not None:
                import torch.distributed as dist
                dist.all_reduce(self.totals, op=dist.ReduceOp.SUM, group=self.group)
            self._decide(self.totals, self.n_total, read=False)
        rcps.loss_table(self.counts, self.px, out=self.table, first_visited_dev=self.result[3:])

    def run(self, after_replay=None):
        """One calibration: replay the graph, wait for the 16-byte result.  ``after_replay`` (optional callable) runs right
        after the replay was enqueued - benchmarks record a CUDA event there, so that the event brackets the device work only."""
        self.graph.replay()
        if after_replay is not None:
            after_replay()
        if self.fused:
            # the kernel stores the result into mapped pinned memory and tags it with its launch epoch: spin on the tag
            self._launches += 1
            if _lib.load().im2im_host_wait_flag(self.result_host.data_ptr() + 16, self._launches, 2_000_000) != 0:
                torch.cuda.current_stream(self.result.device).synchronize()   # slow or failed launch: surface CUDA errors
                if int(self._result_np[4]) != self._launches:
                    raise _lib.Im2ImError("im2im_rcps_calibrate_fused finished without publishing its result")
        else:
            self.result_host[:4].copy_(self.result, non_blocking=True)
            torch.cuda.current_stream(self.result.device).synchronize()
        res = self._result_np
        if int(res[1]) == -3:
            raise _lib.Im2ImError("im2im_rcps_calibrate_fused: a wait between thread blocks of one GPU gave up (blocks not "
                                  "co-resident?); results of this launch are invalid - build a new RcpsGraph")
        if int(res[1]) == -2:
            raise _lib.Im2ImError("a peer rank did not publish its totals within IM2IM_P2P_TIMEOUT_S (crashed or not "
                                  "calling in lockstep); the ranks are out of step - close() this RcpsGraph on every rank "
                                  "and build a new one")
        stop, decided = int(res[0]), bool(res[1])
        lhat = self.lambdas[stop] if stop >= 0 else self.default_lhat
        return lhat, stop, decided

    def tail_stamps_us(self):
        """Profiling aid (fused path): microseconds, relative to the first block's start, at which the LAST launch's last block
        took its ticket, had read the totals, had pushed them to every peer, saw all peers arrive, and published the decision."""
        if not self.fused:
            return None
        torch.cuda.current_stream(self.result.device).synchronize()
        st = self.workspace[64:128].view(torch.int64).cpu().tolist()
        names = ("ticket", "totals_read", "pushed", "peers_arrived", "published")
        return {n: (st[i + 1] - st[0]) / 1e3 for i, n in enumerate(names) if st[i + 1] != 0}

    def close(self):
        """Destroy the captured graph (it may hold NCCL kernels and peer mappings) - call before
        ``dist.destroy_process_group()``: NCCL waits at communicator teardown for graphs that captured its collectives."""
        if self.graph is not None:
            torch.cuda.synchronize(self.result.device)
            self.graph.reset()
            self.graph = None
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    def replay_on_host(self, stats: Optional[dict] = None):
        """Guard-band case: the reference's own expression on the ambiguous columns (sweep.find_stop_index)."""
        px = self.px
        torch.cuda.current_stream(self.result.device).synchronize()   # fused path: run() did not wait for the stream

        def column_to_losses(col):
            return rcps.loss_table(col.reshape(-1, 1).contiguous(), px)[:, 0]

        lhat, stop, visited = sweep.sweep_from_counts(self.counts, self.totals, px, self.config, column_to_losses,
                                                      ascending=True, group=self.group, stats=stats,
                                                      n_total=self.n_total, totals_already_reduced=True)
        rcps.loss_table(self.counts, px, first_visited_col=max(stop, 0), out=self.table)
        return lhat, stop


def rcps_sweep(outputs: torch.Tensor, labels: torch.Tensor, config: dict, device=None, group=None,
               verbose: bool = False, stats: Optional[dict] = None, n_total: Optional[int] = None,
               head: int = _lib.IM2IM_HEAD_QUANTILES):
    """lambda-hat and the loss table from head score planes (N,3|2,C,H,W; see rcps.py) + labels (N,C,H,W), CPU- or
    CUDA-resident; ``head`` is the IM2IM_HEAD_* kind of the planes.

    Returns (lhat 0-dim fp32 CPU tensor, stop index or -1, counts int32 CUDA (N_local, L), visited bool mask (L,)).
    With a torch.distributed ``group`` every rank passes its own contiguous shard of images; the per-lambda totals are
    all-reduced (one int64 vector of length L over NCCL) and every rank reaches the same decision.
    """
    device = _cuda_device(device if device is not None else config['device'])
    lam_dev, order, ascending = _sorted_grid(config, device)
    n_local = outputs.shape[0]
    px = labels[0].numel() if n_local else 0
    counts, totals = _miss_counts_any_device(outputs, labels, lam_dev, device, head=head)
    return _sweep_counts(counts, totals, order, ascending, px, config, group=group, verbose=verbose, stats=stats,
                         n_total=n_total)


def _sorted_grid(config: dict, device):
    """(ascending lambda grid on the device, permutation or None, was it ascending already) - the kernels rank on a sorted
    grid; a descending one (minimum_lambda > maximum_lambda) is un-permuted afterwards."""
    lambdas, dlambda, lam_prime, default_lhat = sweep.lambda_grid(config)
    if not bool(torch.isfinite(lam_prime).all()):
        raise ValueError("lambda grid must be finite")
    ascending = bool((lam_prime[1:] >= lam_prime[:-1]).all())
    if ascending:
        return lam_prime.to(device), None, True
    lam_sorted, order = torch.sort(lam_prime)
    return lam_sorted.to(device), order, False


def _sweep_counts(counts, totals, order, ascending, px, config, group=None, verbose=False, stats=None, n_total=None):
    """The stopping rule on per-image miss counts (N_local, L) + their column totals; returns as ``rcps_sweep``."""
    device = counts.device
    L = counts.shape[1]
    if order is not None:
        inv = torch.empty_like(order)
        inv[order] = torch.arange(L)
        counts = counts[:, inv.to(device)].contiguous()
        totals = totals[inv.to(device)].contiguous()

    def column_to_losses(col: torch.Tensor) -> torch.Tensor:
        return rcps.loss_table(col.reshape(-1, 1).contiguous(), px)[:, 0]

    lhat, stop, visited = sweep.sweep_from_counts(counts, totals, px, config, column_to_losses, ascending=ascending,
                                                  group=group, verbose=verbose, stats=stats, n_total=n_total,
                                                  device_decide=None if verbose else device_decide_fn(px, config))
    return lhat, stop, counts, visited


def calibrate_from_outputs(model, outputs: torch.Tensor, labels: torch.Tensor, config: dict, group=None,
                           table_device: str = "cpu", stats: Optional[dict] = None):
    """Stage 2 of ``calibrate_model`` on precomputed head outputs: sets ``model.lhat`` and returns the loss table.

    The table is (N, L) fp32 with never-visited columns zero (calibrate_model.py:133-136), on the CPU by default like
    the reference's; pass table_device='cuda' to keep it in HBM.
    """
    device = _cuda_device(config['device'])
    with torch.no_grad():
        kind, scores = _head_scores(model, outputs, device)
        lhat, stop, counts, visited = rcps_sweep(scores, labels, config, device=device, group=group, stats=stats,
                                                 head=kind)
        model.set_lhat(lhat)
        return model, _table_from_counts(counts, visited, labels[0].numel(), table_device)


def _table_from_counts(counts, visited, px, table_device: str = "cpu"):
    """(N, L) fp32 loss table with never-visited columns zero (calibrate_model.py:133-136), pinned host memory by default."""
    L = counts.shape[1]
    first = int(torch.nonzero(visited)[0]) if bool(visited.any()) else L
    contiguous_suffix = bool(visited[first:].all())
    table = rcps.loss_table(counts, px, first_visited_col=first if contiguous_suffix else 0)
    if not contiguous_suffix:  # duplicate lambdas: a non-suffix set of columns was written
        table = table * visited.to(device=table.device, dtype=table.dtype)[None, :]
    if table_device == "cpu":
        host = torch.empty(table.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(table, non_blocking=False)
        table = host
    return table


def _tensor_pair(dataset):
    """(inputs, labels) when `dataset` is a two-tensor TensorDataset or a Subset of one over a contiguous range."""
    if isinstance(dataset, TensorDataset) and len(dataset.tensors) == 2:
        return dataset.tensors
    if isinstance(dataset, torch.utils.data.Subset) and isinstance(dataset.dataset, TensorDataset) \
            and len(dataset.dataset.tensors) == 2 and isinstance(dataset.indices, range) and dataset.indices.step == 1:
        r = dataset.indices
        return tuple(t[r.start:r.stop] for t in dataset.dataset.tensors)
    return None


def collect_outputs(model, dataset, config, device):
    """Stage 1 of ``calibrate_model`` (reference :106-123): run the model over the calibration set.

    The reference parks outputs and labels on the CPU and re-uploads them at every lambda step; here they stay in HBM.
    """
    if config['dataset'] == 'temca':
        labels = torch.cat([x[1].unsqueeze(0).to(device) for x in iter(dataset)], dim=0)
        outputs = torch.cat([model(x[0].unsqueeze(0).to(device)) for x in iter(dataset)])
        return outputs, labels
    pair = _tensor_pair(dataset)
    if pair is not None:
        # TensorDataset (or a contiguous Subset of one): slicing the tensors is what the DataLoader's default collate
        # (torch.stack of the items) produces, without the per-item Python loop and the extra host copy
        xs, ys = pair
        n = xs.shape[0]
        bs = int(config['batch_size'])
        if n == 0:
            raise IndexError("empty calibration set")         # the reference fails on dataset[0]
        labels = ys.to(device, non_blocking=True).to(torch.get_default_dtype())   # reference: assigned into torch.zeros(...)
        outputs = None
        for lo in range(0, n, bs):
            out = model(xs[lo:lo + bs].to(device, non_blocking=True))
            if outputs is None:
                outputs = torch.empty((n,) + tuple(out.shape[1:]), dtype=out.dtype, device=device)
            outputs[lo:lo + out.shape[0]] = out
        return outputs, labels
    labels_shape = list(dataset[0][1].unsqueeze(0).shape)
    labels_shape[0] = len(dataset)
    labels = torch.zeros(tuple(labels_shape), device=device)
    outputs_shape = list(model(dataset[0][0].unsqueeze(0).to(device)).shape)
    outputs_shape[0] = len(dataset)
    outputs = torch.zeros(tuple(outputs_shape), device=device)
    loader = DataLoader(dataset, num_workers=0, batch_size=config['batch_size'], pin_memory=True)
    counter = 0
    for batch in loader:
        b = batch[0].shape[0]
        outputs[counter:counter + b] = model(batch[0].to(device, non_blocking=True))
        labels[counter:counter + b] = batch[1].to(device, non_blocking=True)
        counter += b
    return outputs, labels


_HIST_MAX_LAMBDAS = 4096   # the head-fused histogram keeps 16 bytes per lambda in shared memory next to the convolution


def streaming_applicable(model, dataset, config) -> bool:
    """Streaming calibration covers the built-in heads whose scores ARE the head outputs (every head but softmax), on
    map-style datasets; config['streaming_calibration'] = False keeps the two-stage path (outputs materialised)."""
    if not config.get('streaming_calibration', True) or config.get('dataset') == 'temca':
        return False
    fused = _fused_head(model)
    return fused is not None and fused[1] is None


_COPY_STREAMS: dict = {}


def _copy_stream(device) -> "torch.cuda.Stream":
    """One upload stream per device for the life of the process: the caching allocator keeps a pool per stream, so a fresh
    stream per call meant fresh cudaMallocs per call (seen as a 0.2 s outlier on every second calibration)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _COPY_STREAMS.get(key)
    if st is None:
        st = _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def stream_miss_counts(model, dataset, config, device, lam_dev: torch.Tensor, stats: Optional[dict] = None):
    """Run the model over a map-style dataset batch by batch and keep only the per-image miss counts on the ascending grid
    ``lam_dev``: (counts int32 (N, L) on the device, totals int64 (L,), pixels per image).  The (N, 3, C, H, W) output
    tensor of the reference's stage 1 (calibrate_model.py:106-123, eval.py:100-112) is never built.

    With the native bf16 engine and a one-channel quantile head the ranks are booked by the head convolution's own
    epilogue (``UNetInferenceEngine.forward_hist``: the head tensor is never written at all); otherwise the batch's
    outputs live for one ``miss_counts`` pass and are dropped."""
    from ..models.unet_engine import UNetInferenceEngine, native_forward_applicable
    kind, _ = _fused_head(model)
    L = lam_dev.numel()
    n = len(dataset)
    if n == 0:
        raise IndexError("empty calibration set")             # the reference fails on dataset[0]
    bs = int(config['batch_size'])
    pair = _tensor_pair(dataset)
    if pair is not None:
        batches = ((pair[0][lo:lo + bs], pair[1][lo:lo + bs]) for lo in range(0, n, bs))
    else:
        batches = ((b[0], b[1]) for b in DataLoader(dataset, num_workers=0, batch_size=bs, pin_memory=True))
    counts = torch.empty((n, L), dtype=torch.int32, device=device)
    totals = torch.zeros((L,), dtype=torch.int64, device=device)
    hist = None
    px = None
    lo = 0
    fused_batches = 0
    # the next batch's host->device copy runs on its own stream under the current batch's kernels (pinned sources)
    copy_stream = _copy_stream(device)
    main = torch.cuda.current_stream(device)

    def upload(pair_):
        with torch.cuda.stream(copy_stream):
            x_ = pair_[0].to(device, non_blocking=True)
            y_ = pair_[1].to(device, non_blocking=True).to(torch.get_default_dtype()).contiguous()
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return x_, y_, ready

    it = iter(batches)
    nxt = next(it, None)
    pending = upload(nxt) if nxt is not None else None
    while pending is not None:
        xb, yb, ready = pending
        nxt = next(it, None)
        pending = upload(nxt) if nxt is not None else None
        main.wait_event(ready)
        xb.record_stream(main)
        yb.record_stream(main)
        b = xb.shape[0]
        px = yb[0].numel()
        eng = None
        if (kind == _lib.IM2IM_HEAD_QUANTILES and L <= _HIST_MAX_LAMBDAS and yb.dim() == 4 and yb.shape[1] == 1
                and yb.dtype == torch.float32 and native_forward_applicable(model, xb)):
            eng = model.__dict__.get("_native_engine")
            if eng is None:
                eng = model.__dict__["_native_engine"] = UNetInferenceEngine(model)
            if not eng.hist_applicable(xb):
                eng = None
        if eng is not None:
            if hist is None or hist.shape[0] < b:
                hist = torch.zeros((b, L + 1), dtype=torch.int32, device=device)
            eng.forward_hist(xb, yb, lam_dev, hist[:b])
            rcps.counts_from_hist(hist[:b], counts[lo:lo + b], totals)
            fused_batches += 1
        else:
            out = model(xb)
            _count_batch(out, yb, lam_dev, counts[lo:lo + b], totals, kind)
            del out
        lo += b
    if stats is not None:
        stats["streaming"] = True
        stats["head_fused_batches"] = fused_batches
    return counts, totals, px


def calibrate_streaming(model, dataset, config, device, group=None, stats: Optional[dict] = None,
                        table_device: str = "cpu"):
    """Stages 1+2 of ``calibrate_model`` (reference :106-136) without the (N, 3, C, H, W) output tensor: each batch goes
    model -> per-image miss counts at once (``stream_miss_counts``) and only the (N, L) int32 counts stay in HBM.  Counts,
    lhat and the loss table are bit-identical to the two-stage path (same per-pixel rank code, integer sums)."""
    lam_dev, order, ascending = _sorted_grid(config, device)
    counts, totals, px = stream_miss_counts(model, dataset, config, device, lam_dev, stats=stats)
    lhat, stop, counts, visited = _sweep_counts(counts, totals, order, ascending, px, config, group=group, stats=stats)
    model.set_lhat(lhat)
    return model, _table_from_counts(counts, visited, px, table_device)


def _count_batch(out, labels, lam_dev, counts_rows, totals, kind):
    """miss counts of one batch into its rows of the counts table; the running totals keep accumulating."""
    counts_rows.zero_()
    rcps.miss_counts(out, labels, lam_dev, counts=counts_rows, totals=totals, zero=False, head=kind)


def rank_shard(dataset, group):
    """This rank's contiguous block of a map-style calibration set (rank order = row order of the loss table)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = len(dataset)
    if n < world:
        raise ValueError(f"calibration set of {n} images cannot be sharded over {world} ranks")
    lo, hi = n * rank // world, n * (rank + 1) // world
    return torch.utils.data.Subset(dataset, range(lo, hi))


def gather_loss_table(table: torch.Tensor, group, device) -> torch.Tensor:
    """All ranks' rows of the loss table, concatenated in rank order (= image order of the unsharded set), on every rank.
    Off the hot path: the reference's table has one row per calibration image, and a caller that saves it (router.py:138)
    needs all of them."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows = torch.tensor([table.shape[0]], dtype=torch.int64, device=device)
    all_rows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(all_rows, rows, group=group)
    counts = [int(r) for r in all_rows]
    pad = torch.zeros((max(counts), table.shape[1]), dtype=table.dtype, device=device)
    pad[:table.shape[0]] = table.to(device)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0).to(table.device)


def calibrate_model(model, dataset, config, group=None, gather_table: bool = False):
    """Drop-in for the reference's ``calibrate_model``: returns ``(model, calib_loss_table)`` with ``model.lhat`` set.

    ``group`` (optional, a torch.distributed process group, one process per GPU) shards the calibration set: every rank
    runs the model over its contiguous block of a map-style ``dataset`` and keeps its rows of the table; the only
    collective on the data path is the all-reduce of the per-lambda miss totals (int64[L]); all ranks get the same lhat.
    ``gather_table=True`` returns the FULL (N, L) table on every rank, like the reference's, at the cost of one all-gather."""
    with torch.no_grad():
        print(f"Calibrating...")
        model.eval()
        device = _cuda_device(config['device'])
        get_rcps_loss_fn(config)  # raises NotImplementedError for unknown losses, like the reference
        model = model.to(device)
        if group is not None:
            dataset = rank_shard(dataset, group)
        if streaming_applicable(model, dataset, config):
            model, calib_loss_table = calibrate_streaming(model, dataset, config, device, group=group)
        else:
            outputs, labels = collect_outputs(model, dataset, config, device)
            model, calib_loss_table = calibrate_from_outputs(model, outputs, labels, config, group=group)
        if group is not None and gather_table:
            calib_loss_table = gather_loss_table(calib_loss_table, group, device)
        print(f"Model's lhat set to {model.lhat}")
        return model, calib_loss_table
