"""Data-parallel training step: the unit the reference's ``train()`` loops execute
(train_mtmm.py:205-245, train_sd.py:217-282, train.py:186-199) — H2D of the batch, forward, loss
head, backward, gradient all-reduce, SGD step — one process per GPU.

The reference is single-GPU (no DataParallel / distributed code anywhere); the
data-parallel part is new: clips are sharded on the clip axis (a clip is never split: temporal
mixing stays inside a clip), gradients live in ONE flat buffer that is all-reduced bucket by bucket
with NCCL while backward is still running.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


class GradBuckets:
    """Flat gradient storage + overlapped bucketed all-reduce.

    Every parameter's ``.grad`` is a view into one flat fp32 buffer (zeroed with a single memset per
    step).  Parameters are split into ``n_buckets`` contiguous ranges in REVERSE registration order
    (≈ the order backward produces them).  A post-accumulate-grad hook counts finished parameters;
    when a bucket is complete its slice is all-reduced asynchronously on ``comm_stream`` and
    ``finish()`` joins the streams.  With world_size == 1 (or no process group) nothing is sent.
    """

    def __init__(self, params: Sequence[torch.nn.Parameter], n_buckets: int = 3, process_group=None,
                 average: bool = True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBuckets needs at least one trainable parameter")
        dev = self.params[0].device
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.average = average
        order = list(reversed(self.params))          # backward order first -> contiguous slices per bucket

        def pad(n):                                   # every slice on a 32-byte boundary: the chain's kernels
            return (n + 7) // 8 * 8                   # write their weight gradients straight into it

        total = sum(pad(p.numel()) for p in order)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._slices, self._bucket_of, self._offset_of, off = [], {}, {}, 0
        per_bucket = -(-total // max(1, n_buckets))
        bounds: List[List[int]] = []
        for p in order:
            # buckets are opened one after the other: a new one only when the current one is full (a single
            # parameter larger than a bucket's share simply fills its bucket; no bucket index is ever skipped)
            if not bounds or (bounds[-1][1] - bounds[-1][0] >= per_bucket and len(bounds) < n_buckets):
                bounds.append([off, off])
            b = len(bounds) - 1
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[id(p)] = b
            self._offset_of[id(p)] = off
            off += pad(n)
            bounds[b][1] = off
        self.bounds = bounds
        self._need = [0] * len(bounds)
        for p in order:
            self._need[self._bucket_of[id(p)]] += 1
        self._left = list(self._need)
        self._done = set()                   # parameters already counted in this step (a parameter reports once)
        self._works = []
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params] \
            if self.world > 1 else []

    # -- per step -----------------------------------------------------------------------------
    def zero(self):
        self.flat.zero_()
        self._left = list(self._need)
        self._done = set()
        self._works = []
        # a caller's ``optimizer.zero_grad(set_to_none=True)`` or a replaced ``.grad`` would silently detach a
        # parameter from the flat buffer (the fused optimiser and the all-reduce only see the buffer): re-attach
        base = self.flat.data_ptr()
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != base + 4 * self._offset_of[id(p)]:
                off = self._offset_of[id(p)]
                p.grad = self.flat[off:off + p.numel()].view_as(p)

    def _launch(self, b: int):
        lo, hi = self.bounds[b]
        chunk = self.flat[lo:hi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                if self.average:
                    chunk.div_(self.world)
                self._works.append(dist.all_reduce(chunk, group=self.group, async_op=True))
        else:  # CPU / gloo (tests)
            if self.average:
                chunk.div_(self.world)
            self._works.append(dist.all_reduce(chunk, group=self.group, async_op=True))

    def _on_grad(self, p):
        # A parameter whose gradient the fused chain wrote into the flat buffer reports through mark_done(); autograd
        # still runs its AccumulateGrad node afterwards (with an undefined gradient) and fires the post-accumulate hook a
        # second time.  Counting it twice let a bucket that mixes several autograd Functions (the SD exit heads + the
        # end of the backbone) start its all-reduce before its last gradients were written: replicas diverged
        # (found by ranks_in_sync() on the 2-GPU SD run of round 2).
        if id(p) in self._done:
            return
        self._done.add(id(p))
        b = self._bucket_of[id(p)]
        self._left[b] -= 1
        if self._left[b] == 0 and self.world > 1:
            self._launch(b)

    # -- gradient sink protocol (fused.grad_sink): the fused chain writes into the flat buffer itself ------
    def view_for(self, p):
        """The zeroed flat-buffer slice for parameter p, or None if p is not managed here."""
        if id(p) not in self._bucket_of or p.grad is None:
            return None
        g = p.grad
        lo = self.flat.data_ptr()
        if g.dtype != torch.float32 or not (lo <= g.data_ptr() < lo + self.flat.numel() * 4) or not g.is_contiguous():
            return None                      # someone replaced .grad: fall back to autograd accumulation
        return g

    def mark_done(self, params):
        for p in params:
            self._on_grad(p)

    def finish(self):
        """Call after backward: flush buckets whose hooks did not all fire (unused parameters) and
        make the current stream wait for the reductions."""
        if self.world > 1:
            for b, left in enumerate(self._left):
                if left > 0:
                    self._left[b] = 0
                    self._launch(b)
            for w in self._works:
                w.wait()
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._works = []


def build_sgd(model, lr: float, momentum: float = 0.9, weight_decay: float = 5e-4):
    """SGD over the model's nine policy groups with lr_mult / decay_mult applied
    (train_mtmm.py:576-585)."""
    policies = [g for g in model.get_optim_policies() if len(g['params']) > 0]
    for g in policies:
        g['lr'] = lr * g['lr_mult']
        g['weight_decay'] = weight_decay * g['decay_mult']
    fused = all(p.is_cuda for g in policies for p in g['params'])
    return torch.optim.SGD(policies, momentum=momentum, **({"fused": True} if fused else {}))


class FlatSGD:
    """torch.optim.SGD(momentum, weight_decay) over the policy groups of ``get_optim_policies``
    (train_mtmm.py:576-585) as ONE kernel (csrc/optim.cu) over flat buffers laid out like ``GradBuckets.flat``:
    every parameter's storage is moved into ``flat_p`` (``p.data`` becomes a view), the momentum lives in
    ``flat_m`` (zero-initialised = torch's first-step rule).  The base learning rate is a DEVICE scalar:
    ``set_base_lr`` (or ``adjust_learning_rate``) rewrites it without invalidating a captured CUDA graph."""

    def __init__(self, policies, buckets: "GradBuckets", lr: float, momentum: float = 0.9, weight_decay: float = 5e-4):
        self.buckets = buckets
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        dev = buckets.flat.device
        groups = [g for g in policies if len(g['params']) > 0]
        if len(groups) > 64:
            raise ValueError("FlatSGD supports at most 64 parameter groups")
        self.param_groups = [{'params': list(g['params']), 'lr_mult': float(g.get('lr_mult', 1.0)),
                              'decay_mult': float(g.get('decay_mult', 1.0)), 'name': g.get('name', ''),
                              'lr': lr * float(g.get('lr_mult', 1.0))} for g in groups]
        self.flat_p = torch.zeros_like(buckets.flat)
        self.flat_m = torch.zeros_like(buckets.flat)
        code = torch.full((buckets.flat.numel(),), 255, dtype=torch.uint8)
        with torch.no_grad():
            for k, g in enumerate(self.param_groups):
                for p in g['params']:
                    off = buckets._offset_of.get(id(p))
                    if off is None:              # frozen parameter: not managed, never updated
                        continue
                    n = p.numel()
                    self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
                    p.data = self.flat_p[off:off + n].view_as(p)
                    code[off:off + n] = k
        self.code = code.to(dev)
        self.lr_mult = torch.tensor([g['lr_mult'] for g in self.param_groups], dtype=torch.float32, device=dev)
        self.decay_mult = torch.tensor([g['decay_mult'] for g in self.param_groups], dtype=torch.float32, device=dev)
        self.lr_dev = torch.tensor([float(lr)], dtype=torch.float32, device=dev)
        # bf16 mirror of every parameter (same offsets), kept current by the SGD kernel itself: the tensor-core GEMMs
        # stage their weights from it (fused.weight_mirrors) instead of casting 34 weights per step
        self.flat_p16 = self.flat_p.to(torch.bfloat16) if dev.type == "cuda" else None
        self._mirror_of = {}
        if self.flat_p16 is not None:
            for g in self.param_groups:
                for p in g['params']:
                    off = buckets._offset_of.get(id(p))
                    if off is not None:
                        self._mirror_of[id(p)] = self.flat_p16[off:off + p.numel()].view(p.shape)
        self._host_lr = torch.empty(1, dtype=torch.float32).pin_memory() if dev.type == "cuda" else None

    def set_base_lr(self, lr: float):
        for g in self.param_groups:
            g['lr'] = lr * g['lr_mult']
        if self._host_lr is not None:
            self._host_lr[0] = float(lr)
            self.lr_dev.copy_(self._host_lr, non_blocking=True)
        else:
            self.lr_dev.fill_(float(lr))

    def refresh_mirror(self):
        """Re-derive the bf16 mirror from the fp32 parameters (after load_state_dict / a broadcast / any outside write)."""
        if self.flat_p16 is not None:
            self.flat_p16.copy_(self.flat_p)

    def mirror_for(self, p):
        """bf16 view of parameter p inside the mirror, or None (fused.weight_mirrors protocol)."""
        return self._mirror_of.get(id(p))

    def step(self, ema: Optional[torch.Tensor] = None, ema_decay: float = 0.0):
        """One optimiser step; with ``ema`` (a flat fp32 tensor laid out like ``flat_p``) the EMA of the parameters is
        updated in the same kernel from the freshly stepped values (FlatEMA)."""
        from . import _lib
        f = self.buckets.flat
        _lib.call("ehgr_sgd_step", self.flat_p.data_ptr(), f.data_ptr(), self.flat_m.data_ptr(), self.code.data_ptr(),
                  self.lr_mult.data_ptr(), self.decay_mult.data_ptr(), len(self.param_groups), self.lr_dev.data_ptr(),
                  self.momentum, self.weight_decay, f.numel(), _lib.ptr(ema), float(ema_decay), _lib.ptr(self.flat_p16),
                  _lib.stream_ptr(f.device), algo_bytes=f.numel() * (23 + (8 if ema is not None else 0)))

    def zero_grad(self, set_to_none: bool = False):
        self.buckets.zero()


class FlatEMA:
    """The reference's ``EMAWrapper`` (train_mtmm.py:110-140): an exponential moving average of EVERY ``state_dict``
    entry, updated after each optimiser step (train_mtmm.py:242-245).  Same surface (``update``, ``set``, ``forward``,
    ``state_dict``, ``load_state_dict``, ``.model``, ``.decay``); instead of one tiny op group per tensor (444 entries
    for TSM-MobileNetV2 MTMM) the update is three launches:
      * parameters     — inside ``ehgr_sgd_step`` (``FlatSGD.step(ema=...)``): the EMA copy lives in ``flat`` with the
                         layout of ``FlatSGD.flat_p`` and the EMA model's parameters are views into it;
      * float buffers  — BatchNorm running statistics of model and EMA model are moved into two flat tensors (the
                         modules keep views), one ``ehgr_ema_update`` launch;
      * int64 counters — ``num_batches_tracked``, likewise (evaluated in fp32 and truncated, as the reference's
                         ``decay * e + (1. - decay) * m`` followed by ``copy_`` does).
    Results are bit-identical to the reference expression."""

    def __init__(self, model, opt: "FlatSGD", decay: float = 0.9999):
        from copy import deepcopy
        self.decay = float(decay)
        self.opt = opt
        self.model = deepcopy(model)
        for p in self.model.parameters():
            p.requires_grad_(False)
        dev = opt.flat_p.device
        self.flat = opt.flat_p.detach().clone()
        extra_f = []                                     # (training tensor, ema tensor): float entries outside flat_p
        for pm, pe in zip(model.parameters(), self.model.parameters()):
            off = opt.buckets._offset_of.get(id(pm))
            if off is None:                              # frozen parameter: goes with the buffers
                extra_f.append((pm, pe))
            else:
                pe.data = self.flat[off:off + pm.numel()].view_as(pm)
        ints = []
        for bm, be in zip(model.buffers(), self.model.buffers()):
            (extra_f if bm.is_floating_point() else ints).append((bm, be))

        def flatten(pairs, dtype):
            n = sum(a.numel() for a, _ in pairs)
            src, dst = torch.zeros(n, dtype=dtype, device=dev), torch.zeros(n, dtype=dtype, device=dev)
            off = 0
            with torch.no_grad():
                for a, b in pairs:
                    k = a.numel()
                    src[off:off + k].copy_(a.detach().reshape(-1).to(dtype))
                    dst[off:off + k].copy_(b.detach().reshape(-1).to(dtype))
                    a.data = src[off:off + k].view(a.shape)      # the modules keep views: kernels write running statistics here
                    b.data = dst[off:off + k].view(b.shape)
                    off += k
            return src, dst
        if any(a.dtype != torch.float32 for a, _ in extra_f) or any(a.dtype != torch.int64 for a, _ in ints):
            raise TypeError("FlatEMA expects float32 buffers and int64 counters")
        self.buf_src, self.buf_ema = flatten(extra_f, torch.float32)
        self.int_src, self.int_ema = flatten(ints, torch.int64)

    # -- the three launches of one update; the parameter part normally rides on the optimiser step ----------
    def update_buffers(self):
        from . import _lib
        sp = _lib.stream_ptr(self.flat.device)
        if self.buf_src.numel():
            _lib.call("ehgr_ema_update", self.buf_ema.data_ptr(), self.buf_src.data_ptr(), self.buf_src.numel(), self.decay, 0, sp,
                      algo_bytes=12 * self.buf_src.numel())
        if self.int_src.numel():
            _lib.call("ehgr_ema_update", self.int_ema.data_ptr(), self.int_src.data_ptr(), self.int_src.numel(), self.decay, 1, sp)

    def update(self, model=None):
        """EMAWrapper.update for callers that step the optimiser themselves (``FlatSGD.step()`` without ``ema``)."""
        from . import _lib
        _lib.call("ehgr_ema_update", self.flat.data_ptr(), self.opt.flat_p.data_ptr(), self.flat.numel(), self.decay, 0,
                  _lib.stream_ptr(self.flat.device), algo_bytes=12 * self.flat.numel())
        self.update_buffers()

    def set(self, model=None):
        with torch.no_grad():
            self.flat.copy_(self.opt.flat_p)
            self.buf_ema.copy_(self.buf_src)
            self.int_ema.copy_(self.int_src)

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    __call__ = forward

    def state_dict(self):
        return self.model.state_dict()

    def load_state_dict(self, state_dict):
        self.model.load_state_dict(state_dict)       # copy_ into the views: the flat layout is preserved


def adjust_learning_rate(learning_rate, optimizer, epoch, lr_steps):
    """utils.py:39-46.  With ``FlatSGD`` the new rate is written to its device scalar (a captured CUDA graph stays
    valid); with a torch optimiser a train step running in CUDA-graph mode must be told: ``invalidate_graph()``."""
    gamma = 0.1 ** sum(epoch >= s for s in lr_steps)
    if hasattr(optimizer, "set_base_lr"):
        optimizer.set_base_lr(learning_rate * gamma)
        return
    for g in optimizer.param_groups:
        g['lr'] = learning_rate * gamma * g['lr_mult']


def normalize_u8(frames: torch.Tensor, mean=None, std=None, out_dtype=torch.float32) -> torch.Tensor:
    """uint8 frames [..., C, H, W] on the GPU -> (x / 255 - mean[c]) / std[c], the device end of the
    reference's ToTorchFormatTensor(div=True) + GroupNormalize (models/spatial_transforms.py:66-80,489-503);
    bit-identical to those CPU ops for fp32 output.  mean / std: sequences of C floats or None (x / 255 only)."""
    from . import _lib
    _lib.require_cuda(frames)
    if frames.dtype != torch.uint8:
        raise TypeError("normalize_u8 expects uint8 frames")
    frames = frames.contiguous()
    c, h, w = frames.shape[-3:]
    out = torch.empty(frames.shape, dtype=out_dtype, device=frames.device)
    m = torch.tensor(list(mean), dtype=torch.float32, device=frames.device) if mean is not None else None
    sd = torch.tensor(list(std), dtype=torch.float32, device=frames.device) if std is not None else None
    _lib.call("ehgr_normalize_u8", frames.data_ptr(), out.data_ptr(), frames.numel() // (h * w), c, h * w, _lib.ptr(m),
              _lib.ptr(sd), 255.0, _lib.dtype_code(out), _lib.stream_ptr(frames.device),
              algo_bytes=frames.numel() * (1 + out.element_size()))
    return out


def _normalize_with(frames, mean_t, std_t):
    from . import _lib
    frames = frames.contiguous()
    c, h, w = frames.shape[-3:]
    out = torch.empty(frames.shape, dtype=torch.float32, device=frames.device)
    _lib.call("ehgr_normalize_u8", frames.data_ptr(), out.data_ptr(), frames.numel() // (h * w), c, h * w, _lib.ptr(mean_t),
              _lib.ptr(std_t), 255.0, _lib.F32, _lib.stream_ptr(frames.device), algo_bytes=frames.numel() * 5)
    return out


class MTMMTrainStep:
    """One MTMM stage-1 step (train_mtmm.py:205-245) on this rank's shard of clips.

    ``run(rgb, depth, labels)`` takes DEVICE tensors; ``__call__`` takes HOST (pinned) tensors, copies
    them and returns the loss as a Python float (one D2H read), i.e. the end-to-end user call.
    """

    def __init__(self, model, lr=0.00125, momentum=0.9, weight_decay=5e-4, compute_dtype=torch.bfloat16,
                 n_buckets=3, process_group=None, use_graph=False, graph_warmup=2, ema_decay=None):
        from . import fused
        self.model = model
        self.compute_dtype = compute_dtype
        self.device = next(model.parameters()).device
        self.buckets = GradBuckets(list(model.parameters()), n_buckets, process_group)
        if self.device.type == "cuda":   # one fused kernel over flat parameter / gradient / momentum buffers
            self.opt = FlatSGD(model.get_optim_policies(), self.buckets, lr, momentum, weight_decay)
        else:
            self.opt = build_sgd(model, lr, momentum, weight_decay)
        self._fused = fused
        # CUDA-graph mode: the step launches ~3600 small kernels; replaying one captured graph removes the
        # host launch cost.  The first `graph_warmup` calls run eagerly (lazy state: momentum buffers, cuDNN
        # plans, the allocator), the next call is captured with ITS batch in the static input buffers and
        # replayed, so no training step is spent on warm-up data.
        self.use_graph = bool(use_graph) and self.device.type == "cuda"
        self.graph_warmup = graph_warmup
        self._graph = None
        self._static_in = None
        self._static_loss = None
        self._eager_calls = 0
        self._side = None
        self._mean_std = None
        if self.device.type == "cuda" and hasattr(model, "input_mean"):
            self._mean_std = (torch.tensor(model.input_mean, dtype=torch.float32, device=self.device),
                              torch.tensor(model.input_std, dtype=torch.float32, device=self.device))
        self.launches_per_step = None      # libehgr_b200 kernel launches of one step (counted while capturing)
        # model_ema = EMAWrapper(model) of the reference's main() (train_mtmm.py:569), updated after every step (:245)
        self.ema = None
        if ema_decay is not None:
            if not isinstance(self.opt, FlatSGD):
                raise RuntimeError("the fused EMA needs the CUDA optimiser (FlatSGD)")
            self.ema = FlatEMA(model, self.opt, ema_decay)
        self.group = process_group
        if self.buckets.world > 1:
            self.sync_initial_state()

    # -- replica consistency ----------------------------------------------------------------------
    def _src(self):
        return dist.get_global_rank(self.group, 0) if self.group is not None else 0

    def sync_initial_state(self):
        """Every rank starts from rank 0's parameters, momentum and buffers (BatchNorm running statistics,
        num_batches_tracked): ranks that seeded differently or loaded different checkpoints would otherwise
        average gradients taken at different weights and drift apart silently.  One broadcast of the flat
        parameter / momentum buffers (FlatSGD has already moved every trainable parameter into them) plus one
        per unmanaged tensor."""
        with torch.no_grad():
            if isinstance(self.opt, FlatSGD):
                dist.broadcast(self.opt.flat_p, self._src(), group=self.group)
                dist.broadcast(self.opt.flat_m, self._src(), group=self.group)
                managed = {id(p) for p in self.buckets.params}
                rest = [p.data for p in self.model.parameters() if id(p) not in managed]
            else:
                rest = [p.data for p in self.model.parameters()]
            for t in rest + [b.data for b in self.model.buffers()]:
                dist.broadcast(t, self._src(), group=self.group)
            if isinstance(self.opt, FlatSGD):
                self.opt.refresh_mirror()
            if self.ema is not None:
                self.ema.set()

    def sync_bn_buffers(self):
        """BatchNorm batch statistics are per rank while training (the reference has no SyncBN and is
        single-GPU); the running statistics therefore differ slightly between ranks.  Call this before saving a
        checkpoint / evaluating: floating-point buffers are averaged over the ranks, integer ones
        (num_batches_tracked) are taken from rank 0."""
        if self.buckets.world <= 1:
            return
        with torch.no_grad():
            for b in self.model.buffers():
                if b.is_floating_point():
                    dist.all_reduce(b.data, group=self.group)
                    b.data.div_(self.buckets.world)
                else:
                    dist.broadcast(b.data, self._src(), group=self.group)

    def param_checksum(self) -> torch.Tensor:
        """[sum, sum of squares] of all parameters in fp64 — equal on every rank iff the replicas are in sync."""
        with torch.no_grad():
            ps = [self.opt.flat_p] if isinstance(self.opt, FlatSGD) else [p.data for p in self.model.parameters()]
            s = torch.zeros(2, dtype=torch.float64, device=self.device)
            for p in ps:
                d = p.double()
                s[0] += d.sum()
                s[1] += (d * d).sum()
        return s

    def ranks_in_sync(self) -> bool:
        """All-gather the parameter checksum; True iff every rank holds bit-identical sums."""
        if self.buckets.world <= 1:
            return True
        mine = self.param_checksum()
        allc = [torch.empty_like(mine) for _ in range(self.buckets.world)]
        dist.all_gather(allc, mine, group=self.group)
        return all(torch.equal(c, allc[0]) for c in allc)

    def stage(self, rgb_h, depth_h, labels_h):
        return (rgb_h.to(self.device, non_blocking=True), depth_h.to(self.device, non_blocking=True),
                labels_h.to(self.device, non_blocking=True))

    def _prepare(self, rgb, depth=None):
        """uint8 frames (the loader's native format) are normalised on the device; float input passes through."""
        if rgb.dtype == torch.uint8:
            rgb = _normalize_with(rgb, *self._mean_std)
        if depth is not None and depth.dtype == torch.uint8:
            depth = _normalize_with(depth, None, None)
        return rgb, depth

    def _step(self, rgb, depth, labels):
        from .losses import mtmm_loss
        rgb, depth = self._prepare(rgb, depth)
        self.buckets.zero()
        with self._fused.compute_dtype(self.compute_dtype), self._fused.weight_mirrors(self._mirrors()):
            logits, dpred = self.model(rgb)
            loss, _ = mtmm_loss(logits, labels, dpred, depth)
        with self._fused.grad_sink(self.buckets), self._fused.weight_mirrors(self._mirrors()):
            loss.backward()
        self.buckets.finish()
        self._optimizer_step()
        return loss.detach()

    def _mirrors(self):
        return self.opt if isinstance(self.opt, FlatSGD) else None

    def _optimizer_step(self):
        if self.ema is not None:
            self.opt.step(self.ema.flat, self.ema.decay)
            self.ema.update_buffers()
        else:
            self.opt.step()

    def invalidate_graph(self):
        """Call after anything the captured graph has baked in changes (learning rate, train/eval mode,
        frozen parameters, input shapes) and after writing parameters from outside (load_state_dict)."""
        if isinstance(self.opt, FlatSGD):
            self.opt.refresh_mirror()
        self._graph = None
        self._static_in = None
        self._static_loss = None

    def _run_on_side(self, batch):
        """Every step — eager, warm-up or captured — runs on ONE private stream.  Autograd ties the gradient accumulation
        of a leaf to the stream its accumulator was first used on; a step issued from another stream would let the bucket
        all-reduce (ordered after the CALLER's stream) overtake that accumulation: the late local gradient then lands on
        top of the reduced one and the replicas drift apart (caught by ranks_in_sync() in round 2)."""
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            loss = self._step(*batch)
        for t in batch:
            t.record_stream(self._side)
        cur.wait_stream(self._side)
        return loss

    def run(self, *batch):
        if not self.use_graph:
            return self._run_on_side(batch) if self.device.type == "cuda" else self._step(*batch)
        if self._graph is not None and any(a.shape != b.shape or a.dtype != b.dtype
                                           for a, b in zip(batch, self._static_in)):
            self.invalidate_graph()
            self._eager_calls = 0          # new shapes / dtypes: warm up eagerly again before re-capturing
        if self._graph is None:
            # Warm-up calls and the capture run on the same private stream (the documented whole-network recipe).
            cur = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            if self._eager_calls < self.graph_warmup:
                self._eager_calls += 1
                return self._run_on_side(batch)
            from . import _lib
            self._static_in = [torch.empty_like(t) for t in batch]
            for dst, src in zip(self._static_in, batch):
                dst.copy_(src)
            cur.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(graph, stream=self._side):
                self._static_loss = self._step(*self._static_in)
            self.launches_per_step = _lib.launch_count() - l0
            self._graph = graph
        else:
            for dst, src in zip(self._static_in, batch):
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._static_loss

    def __call__(self, rgb_h, depth_h, labels_h) -> float:
        return float(self.run(*self.stage(rgb_h, depth_h, labels_h)).item())


class SDTrainStep(MTMMTrainStep):
    """One self-distillation step (train_sd.py:217-282): forward with the three exit heads, the fused
    SD loss (4 CE + 3 KD + 3 feature terms), backward, gradient all-reduce, SGD.

    The feature term is a SUM over the local batch (train_sd.py:191-193): with gradient averaging over
    `world` ranks its weight would shrink by 1/world relative to a single-process global batch, so
    beta is multiplied by world here (DDP loss scaling)."""

    def __init__(self, model, alpha=0.1, beta=1e-6, temperature=3.0, **kw):
        super().__init__(model, **kw)
        self.alpha, self.beta, self.temperature = alpha, beta * self.buckets.world, temperature

    def stage(self, rgb_h, labels_h):
        return rgb_h.to(self.device, non_blocking=True), labels_h.to(self.device, non_blocking=True)

    def _step(self, rgb, labels):
        from .losses import sd_loss
        rgb, _ = self._prepare(rgb)
        self.buckets.zero()
        with self._fused.compute_dtype(self.compute_dtype), self._fused.weight_mirrors(self._mirrors()):
            outs = self.model(rgb)
            total, _terms = sd_loss(outs[:4], outs[4:], labels, self.alpha, self.beta, self.temperature)
        with self._fused.grad_sink(self.buckets), self._fused.weight_mirrors(self._mirrors()):
            total.backward()
        self.buckets.finish()
        self._optimizer_step()
        return total.detach()

    def __call__(self, rgb_h, labels_h) -> float:
        return float(self.run(*self.stage(rgb_h, labels_h)).item())


class MTMMSDTrainStep(MTMMTrainStep):
    """One step of the combined stage (train_mtmm_sd.py:213-320): forward with the three exit heads and the two depth
    decoders (tsn_mtmm_sd.TSN, modal='rgb_depth'), total = (1-alpha)(CE + 0.01 MSE_depth + 3 CE) + alpha 3 KD + beta 3
    feature, backward of ``total`` (the reference script back-propagates ``loss`` alone, :310 — SURVEY section 4 lists it
    among the script's defects; the quantity it logs and means to minimise is ``total_loss``), all-reduce, SGD."""

    def __init__(self, model, alpha=0.1, beta=1e-6, temperature=3.0, **kw):
        super().__init__(model, **kw)
        self.alpha, self.beta, self.temperature = alpha, beta * self.buckets.world, temperature

    def _step(self, rgb, depth, labels):
        from .losses import mtmm_sd_loss
        rgb, depth = self._prepare(rgb, depth)
        self.buckets.zero()
        with self._fused.compute_dtype(self.compute_dtype), self._fused.weight_mirrors(self._mirrors()):
            outs = self.model(rgb)
            total, _terms, _mse = mtmm_sd_loss(outs[:4], outs[4:8], outs[9], depth, labels, self.alpha, self.beta,
                                               self.temperature)
        with self._fused.grad_sink(self.buckets), self._fused.weight_mirrors(self._mirrors()):
            total.backward()
        self.buckets.finish()
        self._optimizer_step()
        return total.detach()
