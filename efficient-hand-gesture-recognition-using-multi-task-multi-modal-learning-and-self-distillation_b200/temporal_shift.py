"""Temporal shift (TSM) — drop-in for the reference's ``models/temporal_shift.py``.

Same public surface (``TemporalShift``, ``TemporalShift.shift``, ``InplaceShift``, ``TemporalPool``,
``make_temporal_shift``, ``make_temporal_pool``), but the shift itself is one hand-written sm_100a
gather kernel (``csrc/shift.cu``) for forward and its mirror for backward instead of a memset plus
three strided copies (reference: models/temporal_shift.py:27-46) and ~6 autograd copy kernels.

Differences from the reference, on purpose:
  * ``inplace=True`` works (reference raises NotImplementedError, models/temporal_shift.py:34-37,
    with the comment "May need to write a CUDA kernel"): the kernel is out-of-place and race-free,
    and the result is a fresh tensor, which is all any caller uses.
  * ``make_temporal_shift`` also accepts a MobileNetV2 (reference: ResNet only,
    models/temporal_shift.py:111,145-146) and wraps ``conv[0]`` of the ten residual
    InvertedResidual blocks — the predicate of models/models.py:183.
  * channels-last (NHWC) inputs are shifted in NHWC without a layout change.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def _layout_of(x: torch.Tensor):
    """(tensor, layout code): NCHW-contiguous or channels-last; anything else is made contiguous."""
    if x.is_contiguous():
        return x, _lib.NCHW
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last):
        return x, _lib.NHWC
    return x.contiguous(), _lib.NCHW


def _run_shift(x: torch.Tensor, n_segment: int, fold: int, backward: bool) -> torch.Tensor:
    _lib.require_cuda(x)
    nt, c, h, w = x.shape
    x, layout = _layout_of(x)
    out = torch.empty_like(x)  # preserves the memory format
    if x.numel() == 0:
        return out
    name = "ehgr_temporal_shift_bwd" if backward else "ehgr_temporal_shift_fwd"
    with torch.cuda.device(x.device):
        _lib.call(name, x.data_ptr(), out.data_ptr(), nt // n_segment, n_segment, c, h * w, fold,
                  _lib.dtype_code(x), layout, _lib.stream_ptr(x.device),
                  algo_bytes=2 * x.numel() * x.element_size())
    return out


class _ShiftFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_segment, fold):
        ctx.n_segment, ctx.fold = n_segment, fold
        return _run_shift(x, n_segment, fold, backward=False)

    @staticmethod
    def backward(ctx, grad_out):
        return _run_shift(grad_out, ctx.n_segment, ctx.fold, backward=True), None, None


def temporal_shift(x: torch.Tensor, n_segment: int, fold_div: int = 3) -> torch.Tensor:
    """Functional form; ``x`` is ``[n_batch*n_segment, c, h, w]``."""
    if x.dim() != 4:
        raise RuntimeError(f"temporal shift expects a 4-D [nt, c, h, w] tensor, got {tuple(x.shape)}")
    nt, c, h, w = x.size()
    n_batch = nt // n_segment
    if n_batch * n_segment != nt:
        # the reference fails here in x.view(n_batch, n_segment, c, h, w)
        raise RuntimeError(
            f"shape '[{n_batch}, {n_segment}, {c}, {h}, {w}]' is invalid for input of size {x.numel()}")
    fold = c // fold_div
    return _ShiftFunction.apply(x, n_segment, fold)


class TemporalShift(nn.Module):
    """``net(shift(x))`` — reference models/temporal_shift.py:11-25 (same attributes)."""

    def __init__(self, net, n_segment=3, n_div=8, inplace=False):
        super().__init__()
        self.net = net
        self.n_segment = n_segment
        self.fold_div = n_div
        self.inplace = inplace
        if inplace:
            print('=> Using in-place shift...')
        print('=> Using fold div: {}'.format(self.fold_div))

    def forward(self, x):
        x = self.shift(x, self.n_segment, fold_div=self.fold_div, inplace=self.inplace)
        return self.net(x)

    @staticmethod
    def shift(x, n_segment, fold_div=3, inplace=False):
        # `inplace` selects nothing: the kernel is out-of-place and race-free (see module docstring)
        return temporal_shift(x, n_segment, fold_div)


class InplaceShift(torch.autograd.Function):
    """Same contract as the reference's InplaceShift (models/temporal_shift.py:49-76):
    ``input`` is ``[n, t, c, h, w]`` and ``fold`` an absolute channel count; the shifted values are
    written back into ``input``'s storage and ``input`` is returned."""

    @staticmethod
    def forward(ctx, input, fold):
        ctx.fold_ = fold
        n, t, c, h, w = input.size()
        shifted = _run_shift(input.detach().reshape(n * t, c, h, w), t, fold, backward=False)
        input.data.copy_(shifted.view(n, t, c, h, w))
        ctx.mark_dirty(input)
        return input

    @staticmethod
    def backward(ctx, grad_output):
        n, t, c, h, w = grad_output.size()
        g = _run_shift(grad_output.reshape(n * t, c, h, w), t, ctx.fold_, backward=True)
        return g.view(n, t, c, h, w), None


class _TemporalPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_segment):
        _lib.require_cuda(x)
        nt, c, h, w = x.shape
        x, _layout = _layout_of(x)                 # frames are contiguous in both layouts: the kernel sees [n, T, c*h*w]
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        n, t_out = nt // n_segment, (n_segment - 1) // 2 + 1
        out = torch.empty((n * t_out, c, h, w), dtype=x.dtype, device=x.device,
                          memory_format=torch.channels_last if _layout == _lib.NHWC else torch.contiguous_format)
        if x.numel():
            with torch.cuda.device(x.device):
                _lib.call("ehgr_temporal_pool_fwd", x.data_ptr(), out.data_ptr(), n, n_segment, c * h * w, _lib.dtype_code(x),
                          _lib.stream_ptr(x.device), algo_bytes=(x.numel() + out.numel()) * x.element_size())
        ctx.save_for_backward(x)
        ctx.n_segment = n_segment
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        nt, c, h, w = x.shape
        g = g.contiguous(memory_format=torch.channels_last) if (x.dim() == 4 and not x.is_contiguous()) else g.contiguous()
        g = g.to(x.dtype)
        dx = torch.empty_like(x)
        if x.numel():
            with torch.cuda.device(x.device):
                _lib.call("ehgr_temporal_pool_bwd", x.data_ptr(), g.data_ptr(), dx.data_ptr(), nt // ctx.n_segment, ctx.n_segment,
                          c * h * w, _lib.dtype_code(x), _lib.stream_ptr(x.device),
                          algo_bytes=(2 * x.numel() + g.numel()) * x.element_size())
        return dx, None


class TemporalPool(nn.Module):
    """Temporal max-pool k=3,s=2 between stages (reference models/temporal_shift.py:79-98): one kernel over
    [n, T, c*h*w] (csrc/tpool.cu) instead of view / transpose / max_pool3d / transpose / contiguous."""

    def __init__(self, net, n_segment):
        super().__init__()
        self.net = net
        self.n_segment = n_segment

    def forward(self, x):
        x = self.temporal_pool(x, n_segment=self.n_segment)
        return self.net(x)

    @staticmethod
    def temporal_pool(x, n_segment):
        nt, c, h, w = x.size()
        if nt % n_segment:
            raise RuntimeError(f"shape '[{nt // n_segment}, {n_segment}, {c}, {h}, {w}]' is invalid for input of size {x.numel()}")
        return _TemporalPoolFunction.apply(x, n_segment)


def _is_mobilenet_v2(net) -> bool:
    from .mobilenet_v2 import MobileNetV2
    return isinstance(net, MobileNetV2)


def residual_sites(net):
    """The ten InvertedResidual blocks a temporal module is inserted into (models/models.py:183)."""
    from .mobilenet_v2 import InvertedResidual
    return [m for m in net.modules()
            if isinstance(m, InvertedResidual) and len(m.conv) == 8 and m.use_res_connect]


def make_temporal_shift(net, n_segment, n_div=8, place='blockres', temporal_pool=False):
    """Insert ``TemporalShift`` into a backbone (reference models/temporal_shift.py:101-146)."""
    if temporal_pool:
        n_segment_list = [n_segment, n_segment // 2, n_segment // 2, n_segment // 2]
    else:
        n_segment_list = [n_segment] * 4
    assert n_segment_list[-1] > 0
    print('=> n_segment per stage: {}'.format(n_segment_list))

    import torchvision
    if isinstance(net, torchvision.models.ResNet):
        stages = ['layer1', 'layer2', 'layer3', 'layer4']
        if place == 'block':
            for name, seg in zip(stages, n_segment_list):
                blocks = list(getattr(net, name).children())
                print('=> Processing stage with {} blocks'.format(len(blocks)))
                setattr(net, name, nn.Sequential(
                    *[TemporalShift(b, n_segment=seg, n_div=n_div) for b in blocks]))
        elif 'blockres' in place:
            n_round = 1
            if len(list(net.layer3.children())) >= 23:
                n_round = 2
                print('=> Using n_round {} to insert temporal shift'.format(n_round))
            for name, seg in zip(stages, n_segment_list):
                blocks = list(getattr(net, name).children())
                print('=> Processing stage with {} blocks residual'.format(len(blocks)))
                for i, b in enumerate(blocks):
                    if i % n_round == 0:
                        b.conv1 = TemporalShift(b.conv1, n_segment=seg, n_div=n_div)
                setattr(net, name, nn.Sequential(*blocks))
    elif _is_mobilenet_v2(net):
        if temporal_pool:
            raise NotImplementedError('temporal_pool is defined for ResNet stages only')
        for m in residual_sites(net):
            m.conv[0] = TemporalShift(m.conv[0], n_segment=n_segment, n_div=n_div)
    else:
        raise NotImplementedError(place)


def make_temporal_pool(net, n_segment):
    import torchvision
    if isinstance(net, torchvision.models.ResNet):
        print('=> Injecting nonlocal pooling')
        net.layer2 = TemporalPool(net.layer2, n_segment)
    else:
        raise NotImplementedError
