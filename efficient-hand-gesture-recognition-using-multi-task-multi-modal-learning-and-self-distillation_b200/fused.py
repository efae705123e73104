"""Fused execution of the backbone on CUDA (sm_100a): every arithmetic op of the TSM-MobileNetV2
chain below is a hand-written kernel of libehgr_b200.so reached through the C ABI
(include/ehgr_b200.h).  No torch conv / BN / activation op and no CPU path is involved.

Design:
  * activations live in NHWC ("channels last") in the compute dtype (fp32 or bf16) and are stored
    RAW, i.e. as the convolution output before BatchNorm.  Each conv kernel accumulates the batch
    statistics of its output in its epilogue; the consumer kernel applies BatchNorm(+ReLU6) — or the
    temporal shift — while loading ("row operand", csrc/rowop.cuh).  The reference runs conv, BN and
    ReLU6 as three kernels with an HBM round trip each (archs/mobilenet_v2.py:40-59).
  * backward mirrors this: a layer's gradient kernels read (grad, raw) through a BNBWD row operand
    that evaluates the BatchNorm(+ReLU6) backward on the fly, so dz / dBN tensors never exist.
  * one ``torch.autograd.Function`` (``_ChainFunction``) runs a whole chain of units (stem,
    InvertedResidual blocks, final 1x1 conv); the module tree (``features.{i}.conv.{j}``) is only
    the parameter container and the checkpoint contract.
"""
from __future__ import annotations

import contextlib
import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib, action_ops
from ._lib import BnBwd, BnFin, RowOp

import os as _os

_STATE = {"dtype": torch.float32, "engine": 0, "grad_sink": None, "mirrors": None}

# BatchNorm finalisation inside the producing kernel ("last CTA done", csrc/bnfin.cuh): bit 0 forward, bit 1 backward.
# Measured on B200 inside the captured step (profiles/README.md, round 2): the in-kernel tail (device-scope fence,
# ticket, one CTA finalising) costs MORE than a dependent one-block launch inside a CUDA graph (17.2 vs 16.9 ms per
# step with both bits set), so the default keeps the separate ehgr_bn_finalize / ehgr_bn_bwd_finalize launches.
FUSED_FINALIZE = int(_os.environ.get("EHGR_FUSED_FINALIZE", "0"))


@contextlib.contextmanager
def weight_mirrors(provider):
    """While active, ``provider.mirror_for(weight)`` supplies the bf16 mirror of a pointwise-conv weight (or None) —
    train_step.FlatSGD keeps one flat mirror current inside its SGD kernel, so no per-layer cast runs."""
    old = _STATE["mirrors"]
    _STATE["mirrors"] = provider
    try:
        yield
    finally:
        _STATE["mirrors"] = old


@contextlib.contextmanager
def grad_sink(sink):
    """While active, chain backward writes parameter gradients straight into ``sink``'s buffers
    (``sink.view_for(p)`` -> zero-initialised fp32 tensor shaped like p, or None) instead of returning
    them to autograd, and reports finished parameters with ``sink.mark_done(params)``.  Used by
    train_step.GradBuckets: no per-parameter accumulate kernels, no second gradient buffer.  Every chain
    parameter may be used once per backward while a sink is active (BatchNorm gradients are stored, not added)."""
    old = _STATE["grad_sink"]
    _STATE["grad_sink"] = sink
    try:
        yield
    finally:
        _STATE["grad_sink"] = old


@contextlib.contextmanager
def compute_dtype(dtype):
    """Storage dtype of the activations inside the fused chain (fp32 master weights either way).
    fp32: exact-fp32 CUDA-core GEMMs (parity mode); bf16: tcgen05 tensor-core GEMMs."""
    if dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("compute dtype must be torch.float32 or torch.bfloat16")
    old = _STATE["dtype"]
    _STATE["dtype"] = dtype
    try:
        if dtype == torch.bfloat16:
            with torch.autocast("cuda", dtype=torch.bfloat16):  # library modules around the chain (decoder)
                yield
        else:
            yield
    finally:
        _STATE["dtype"] = old


@contextlib.contextmanager
def gemm_engine(engine: int):
    """0 auto, 1 force fp32 SIMT, 2 force tcgen05 (tests)."""
    old = _STATE["engine"]
    _STATE["engine"] = engine
    try:
        yield
    finally:
        _STATE["engine"] = old


def current_dtype():
    return _STATE["dtype"]


# ------------------------------------------------------------------------------------------------
# row operands
# ------------------------------------------------------------------------------------------------
def _bf16_mirror(w, storage_dtype, cache=None):
    """bf16 copy of a pointwise-conv weight for the tensor-core engine's asynchronous B staging
    (None for fp32 storage / the SIMT engine, which read the fp32 master weights).  `cache` is the
    per-forward dictionary that backward reuses (one cast per weight per step); there is deliberately no
    cross-step cache: CUDA-graph replays update weights without touching their version counters."""
    if storage_dtype != torch.bfloat16 or _STATE["engine"] == 1:
        return None
    if cache is not None and id(w) in cache:
        return cache[id(w)]
    prov = _STATE["mirrors"]
    m = prov.mirror_for(w) if prov is not None else None
    if m is None:
        m = w.detach().to(torch.bfloat16)
    if cache is not None:
        cache[id(w)] = m
    return m


def op_plain(t):
    return RowOp(mode=0, in1=t.data_ptr())


def op_affine(raw, scale, shift, relu6):
    return RowOp(mode=1, relu6=int(relu6), in1=raw.data_ptr(), scale=scale.data_ptr(), shift=shift.data_ptr())


def op_shift(x, n_segment, fold, hw, direction=1):
    return RowOp(mode=2, in1=x.data_ptr(), n_segment=n_segment, fold=fold, hw=hw, shift_dir=direction)


def op_bnbwd(g, raw, ca, cb, cc, scale, shift, relu6):
    return RowOp(mode=3, relu6=int(relu6), in1=g.data_ptr(), in2=raw.data_ptr(), scale=scale.data_ptr(),
                 shift=shift.data_ptr(), ca=ca.data_ptr(), cb=cb.data_ptr(), cc=cc.data_ptr())


def op_conv3(src, scale, shift, act, ho, wo, cin, up):
    """im2col row of a dense 3x3 convolution over `src` ([frames, ho >> up, wo >> up, cin], NHWC) with the producer's
    BatchNorm+activation (scale/shift/act; scale None = plain tensor) and a nearest x2 upsample (up) applied on load."""
    return RowOp(mode=5, relu6=int(act) if scale is not None else 0, in1=src.data_ptr(), scale=_lib.ptr(scale),
                 shift=_lib.ptr(shift), hw=ho * wo, cv_h=ho, cv_w=wo, cv_cin=cin, cv_up=int(up))


def _nhwc_empty(nt, h, w, c, dtype, device):
    """logical [NT,C,H,W] tensor with channels-last strides."""
    return torch.empty((nt, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)


def _is_nhwc_dense(x):
    nt, c, h, w = x.shape
    return x.stride() == (h * w * c, 1, w * c, c)


def _as_nhwc(x, dtype):
    if x.dtype == dtype and _is_nhwc_dense(x):
        return x
    nt, c, h, w = x.shape
    y = _nhwc_empty(nt, h, w, c, dtype, x.device)
    y.copy_(x)
    return y


# ------------------------------------------------------------------------------------------------
# chain description
# ------------------------------------------------------------------------------------------------
@dataclass
class Stage:
    kind: str                      # 'stem' | 'pw' | 'dw' | 'conv3' (dense 3x3, pad 1, stride 1: the depth decoder)
    conv: nn.Conv2d
    bn: Optional[nn.BatchNorm2d]   # None: a bare convolution (the depthwise halves of SepConv, models/models_SD.py:81-101)
    relu6: int                     # activation code after the BatchNorm: 0 none, 1 ReLU6, 2 ReLU
    stride: int = 1
    shift: Optional[Tuple[int, int]] = None   # (n_segment, fold) — TemporalShift on the input (pw only)
    action: Optional[nn.Module] = None        # Action module wrapping this conv (pw only, first stage of a unit)
    up: bool = False                          # conv3 only: the input is read through a nearest x2 upsample

    def n_params(self) -> int:
        return (3 if self.bn is not None else 1) + (10 if self.action is not None else 0)


@dataclass
class Unit:
    stages: List[Stage]
    residual: bool = False
    tap: bool = False              # also return this unit's output from the chain


def _conv_bn_stage(kind, conv, bn, relu6, shift=None):
    if conv.bias is not None:
        raise NotImplementedError("fused chain expects bias-free convolutions")
    if bn is not None and (bn.momentum is None or not bn.track_running_stats or not bn.affine):
        raise NotImplementedError("fused chain expects the default nn.BatchNorm2d configuration")
    return Stage(kind, conv, bn, int(relu6), conv.stride[0], shift)


def unit_of_sepconv(m) -> "Unit":
    """SepConv (models/models_SD.py:81-101): dw3x3(stride) -> pw -> BN -> ReLU -> dw3x3 -> pw -> BN -> ReLU as four
    stages of the fused chain — the depthwise halves are bare convolutions (their raw output is the PLAIN operand of
    the pointwise GEMM), the activations are plain ReLU (activation code 2)."""
    op = m.op
    return Unit([_conv_bn_stage('dw', op[0], None, 0), _conv_bn_stage('pw', op[1], op[2], 2),
                 _conv_bn_stage('dw', op[4], None, 0), _conv_bn_stage('pw', op[5], op[6], 2)])


def unit_of_block(block) -> Unit:
    """InvertedResidual -> Unit (archs/mobilenet_v2.py:28-66)."""
    from .action import Action
    from .temporal_shift import TemporalShift
    conv = block.conv
    stages = []
    k = 0
    if len(conv) == 8:
        first, shift = conv[0], None
        if isinstance(first, TemporalShift):
            shift = (first.n_segment, first.net.in_channels // first.fold_div)
            first = first.net
        action = None
        if isinstance(first, Action):
            action, first = first, first.net
        stages.append(_conv_bn_stage('pw', first, conv[1], 1, shift))
        stages[-1].action = action
        k = 3
    stages.append(_conv_bn_stage('dw', conv[k], conv[k + 1], 1))
    stages.append(_conv_bn_stage('pw', conv[k + 3], conv[k + 4], 0))
    return Unit(stages, residual=block.use_res_connect)


def units_of_backbone(model, taps: Sequence[int] = ()) -> List[Unit]:
    """MobileNetV2.features -> chain; the stem is merged with features[1] (its only consumer) so that
    the stem activation is never materialised."""
    from .mobilenet_v2 import InvertedResidual
    feats = model.features
    units: List[Unit] = []
    stem = _conv_bn_stage('stem', feats[0][0], feats[0][1], 1)
    first = True
    for i in range(1, len(feats) - 1):
        blk = feats[i]
        if not isinstance(blk, InvertedResidual):
            raise NotImplementedError(type(blk))
        u = unit_of_block(blk)
        if first:
            if 0 in taps:
                units.append(Unit([stem], tap=True))
            else:
                u.stages.insert(0, stem)
            first = False
        u.tap = i in taps
        units.append(u)
    last = feats[len(feats) - 1]
    units.append(Unit([_conv_bn_stage('pw', last[0], last[1], 1)], tap=(len(feats) - 1) in taps))
    return units


def _has_action(model) -> bool:
    from .action import Action
    return any(isinstance(m, Action) for m in model.modules())


# ------------------------------------------------------------------------------------------------
# the chain autograd Function
# ------------------------------------------------------------------------------------------------
def _launch_conv_fwd(st: Stage, a_op, a_geom, w, out, stats, dev, x_nchw=None, mirrors=None, fin=None):
    """One convolution (+ batch statistics) and, inside the same launch, the BatchNorm finalisation `fin` (the CTA that
    finishes last writes scale / shift / mean / invstd and updates the running statistics: csrc/bnfin.cuh)."""
    nt, h, wd, cin = a_geom
    cout = st.conv.out_channels
    sp = _lib.stream_ptr(dev)
    stats_p = 0 if stats is None else stats.data_ptr()
    code = _lib.dtype_code(out)
    es = out.element_size()
    fin_p = ctypes.byref(fin) if fin is not None else None
    if st.kind == 'stem':
        _lib.call("ehgr_stem_fwd_bn", x_nchw.data_ptr(), w.data_ptr(), out.data_ptr(), stats_p, nt, h, wd, cout,
                  _lib.dtype_code(x_nchw), code, fin_p, sp,
                  algo_bytes=x_nchw.numel() * x_nchw.element_size() + out.numel() * es,
                  algo_flops=2 * 27 * out.numel())
    elif st.kind == 'pw':
        m = nt * h * wd
        w16 = _bf16_mirror(w, out.dtype, mirrors)
        _lib.call("ehgr_pw_gemm_bn", ctypes.byref(a_op), w.data_ptr(), _lib.ptr(w16), 0, out.data_ptr(), 0, stats_p, m,
                  cin, cout, code, _STATE["engine"], fin_p, sp, algo_bytes=m * (cin + cout) * es + cin * cout * 4,
                  algo_flops=2 * m * cin * cout)
    elif st.kind == 'conv3':
        # implicit GEMM over the im2col operand: M = output pixels, K = 9*cin; w = (fp32 packed or the raw weight as a
        # placeholder, bf16 packed) from _pack_conv3
        m = out.shape[0] * out.shape[2] * out.shape[3]
        w32, w16 = w
        _lib.call("ehgr_pw_gemm_bn", ctypes.byref(a_op), w32.data_ptr(), _lib.ptr(w16), 0, out.data_ptr(), 0, stats_p, m,
                  9 * cin, cout, code, _STATE["engine"], fin_p, sp, tag="[conv3x3]",
                  algo_bytes=(nt * h * wd * cin + m * cout) * es + 9 * cin * cout * es, algo_flops=2 * m * 9 * cin * cout)
    else:
        _lib.call("ehgr_dw_fwd_bn", ctypes.byref(a_op), w.data_ptr(), out.data_ptr(), stats_p, nt, h, wd, cin,
                  st.stride, code, fin_p, sp, algo_bytes=(nt * h * wd * cin + out.numel()) * es + 36 * cin,
                  algo_flops=18 * out.numel())


def _pack_conv3(w, dt, dev):
    """The two GEMM layouts of a dense 3x3 filter (csrc/conv3.cu): forward [cout, 9*cin] and dgrad [cin, 9*cout] (flipped,
    transposed), in the storage dtype the engine reads — bf16 for the tensor-core engine, fp32 for the SIMT engine.
    Returns ((w32, w16) forward, (w32, w16) dgrad); the unused member of a pair is the raw weight (placeholder) / None."""
    cout, cin = w.shape[0], w.shape[1]
    simt = dt == torch.float32 or _STATE["engine"] == 1
    pdt = torch.float32 if simt else torch.bfloat16
    wf = torch.empty(cout * 9 * cin, dtype=pdt, device=dev)
    wd = torch.empty(cout * 9 * cin, dtype=pdt, device=dev)
    _lib.call("ehgr_conv3_pack", w.data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, cin, _lib.dtype_code(wf), _lib.stream_ptr(dev))
    return ((wf, None), (wd, None)) if simt else ((w, wf), (w, wd))


class _ChainFunction(torch.autograd.Function):
    """forward(units, dtype, x, *params): params = per stage (conv.weight, bn.weight, bn.bias)."""

    @staticmethod
    def forward(ctx, units: List[Unit], dt, x, *params):
        dev = x.device
        _lib.require_cuda(x)
        stages = [s for u in units for s in u.stages]
        n_stat = sum(2 * s.conv.out_channels for s in stages if s.bn is not None and s.bn.training)
        stat_arena = torch.zeros(max(n_stat, 1), dtype=torch.float64, device=dev)
        vec_arena = torch.empty(sum(4 * s.conv.out_channels for s in stages if s.bn is not None), dtype=torch.float32, device=dev)
        tickets = torch.zeros(len(stages), dtype=torch.int32, device=dev)    # one "last CTA" counter per layer
        s_off = v_off = k_i = 0
        sp = _lib.stream_ptr(dev)

        mirrors = {}                                     # id(weight) -> bf16 mirror, reused by backward
        first_is_stem = stages[0].kind == 'stem'
        if first_is_stem:
            if x.dim() != 4 or x.shape[1] != 3:
                raise RuntimeError(f"stem expects [NT,3,H,W], got {tuple(x.shape)}")
            x_in = x.contiguous()                       # NCHW, read directly by the stem kernel
            if x_in.dtype not in (torch.float32, torch.bfloat16):
                x_in = x_in.float()
            cur_final = None
            geom = (x_in.shape[0], x_in.shape[2], x_in.shape[3], 3)
        else:
            x_in = _as_nhwc(x, dt)
            cur_final = x_in
            geom = (x_in.shape[0], x_in.shape[2], x_in.shape[3], x_in.shape[1])

        saved_units, outputs, nbt = [], [], []
        p_i = 0
        for u in units:
            unit_in, unit_geom = cur_final, geom
            lazy = None                                  # (raw, scale, shift, relu6) of the previous stage
            recs = []
            for st in u.stages:
                w = params[p_i]
                gamma, beta = (params[p_i + 1], params[p_i + 2]) if st.bn is not None else (None, None)
                act_params = params[p_i + (3 if st.bn is not None else 1):p_i + st.n_params()]
                p_i += st.n_params()
                nt, h, wd, cin = geom
                cout = st.conv.out_channels
                act_state = None
                aux = None
                if st.kind == 'stem':
                    a_op = None
                elif st.kind == 'conv3':
                    ho, wo = (h << 1, wd << 1) if st.up else (h, wd)
                    xu = None
                    if lazy is None and st.up:
                        # the nearest x2 upsample of a finished activation is materialised (a few MB): the convolution
                        # then gathers its operand with TMA boxes, which cannot fold the x2 index map (a lazy
                        # BatchNorm+ReLU producer keeps the fused gather instead)
                        xu = _nhwc_empty(nt, ho, wo, cin, dt, dev)
                        _lib.call("ehgr_upsample2_fwd", cur_final.data_ptr(), xu.data_ptr(), nt, h, wd, cin, _lib.dtype_code(xu), sp,
                                  algo_bytes=5 * cur_final.numel() * cur_final.element_size())
                        a_op = op_conv3(xu, None, None, 0, ho, wo, cin, False)
                    elif lazy is None:
                        a_op = op_conv3(cur_final, None, None, 0, ho, wo, cin, False)
                    else:
                        a_op = op_conv3(lazy[0], lazy[1], lazy[2], lazy[3], ho, wo, cin, st.up)
                    aux = _pack_conv3(w, dt, dev) + (xu,)
                elif st.action is not None:
                    if lazy is not None:
                        raise NotImplementedError("Action is defined on a block input")
                    act_state, a_op = action_ops.forward_gates(st.action, cur_final, act_params, dt)
                elif lazy is None:
                    a_op = (op_shift(cur_final, st.shift[0], st.shift[1], h * wd, 1) if st.shift is not None
                            else op_plain(cur_final))
                else:
                    if st.shift is not None:
                        raise NotImplementedError("temporal shift is defined on a block input")
                    a_op = op_affine(*lazy) if lazy[1] is not None else op_plain(lazy[0])
                if st.kind != 'conv3':
                    ho, wo = ((h - 1) // st.stride + 1, (wd - 1) // st.stride + 1) if st.kind != 'pw' else (h, wd)
                raw = _nhwc_empty(nt, ho, wo, cout, dt, dev)
                w_arg = aux[0] if st.kind == 'conv3' else w
                if st.bn is None:                        # bare convolution: no statistics, no finalisation
                    _launch_conv_fwd(st, a_op, geom, w_arg, raw, None, dev, mirrors=mirrors)
                    k_i += 1
                    recs.append((raw, None, geom, False, None, aux))
                    lazy = (raw, None, None, 0)
                    geom = (nt, ho, wo, cout)
                    continue
                tr = st.bn.training
                stats = None
                if tr:
                    stats = stat_arena[s_off:s_off + 2 * cout]
                    s_off += 2 * cout
                vec = vec_arena[v_off:v_off + 4 * cout].view(4, cout)   # scale, shift, mean, invstd
                v_off += 4 * cout
                fin = BnFin(gamma=gamma.data_ptr(), beta=beta.data_ptr(), running_mean=st.bn.running_mean.data_ptr(),
                            running_var=st.bn.running_var.data_ptr(), scale=vec[0].data_ptr(), shift=vec[1].data_ptr(),
                            mean=vec[2].data_ptr(), invstd=vec[3].data_ptr(), counter=tickets[k_i:].data_ptr(),
                            count=nt * ho * wo, momentum=float(st.bn.momentum), eps=float(st.bn.eps), training=int(tr))
                k_i += 1
                _launch_conv_fwd(st, a_op, geom, w_arg, raw, stats, dev, x_nchw=x_in if st.kind == 'stem' else None,
                                 mirrors=mirrors, fin=fin if FUSED_FINALIZE & 1 else None)
                if not FUSED_FINALIZE & 1:
                    _lib.call("ehgr_bn_finalize", 0 if stats is None else stats.data_ptr(), nt * ho * wo, gamma.data_ptr(),
                              beta.data_ptr(), st.bn.running_mean.data_ptr(), st.bn.running_var.data_ptr(),
                              float(st.bn.momentum), float(st.bn.eps), int(tr), vec[0].data_ptr(), vec[1].data_ptr(),
                              vec[2].data_ptr(), vec[3].data_ptr(), cout, sp)
                if tr and st.bn.num_batches_tracked is not None:
                    nbt.append(st.bn.num_batches_tracked)
                recs.append((raw, vec, geom, tr, act_state, aux))
                lazy = (raw, vec[0], vec[1], st.relu6)
                geom = (nt, ho, wo, cout)
            # materialise the unit output: BN(+activation) (+ residual)
            if lazy[1] is None:
                raise NotImplementedError("a unit must end with a BatchNorm stage")
            nt, h, wd, c = geom
            out = _nhwc_empty(nt, h, wd, c, dt, dev)
            _lib.call("ehgr_row_apply", ctypes.byref(op_affine(*lazy)), unit_in.data_ptr() if u.residual else 0,
                      out.data_ptr(), nt * h * wd, c, _lib.dtype_code(out), sp,
                      algo_bytes=(2 + int(u.residual)) * out.numel() * out.element_size())
            saved_units.append((unit_in, unit_geom, recs))
            cur_final = out
            if u.tap:
                outputs.append(out)
                # the next unit saves its input for backward: keep a detached alias there, not the Function output itself
                # (output -> grad_fn -> ctx -> saved input would be a reference cycle that only the cycle GC frees)
                cur_final = out.detach()
        tap_units = [i for i, u in enumerate(units) if u.tap]
        if not units[-1].tap:
            outputs.append(cur_final)
            tap_units.append(len(units) - 1)
        if nbt:
            torch._foreach_add_(nbt, 1)
        ctx.units, ctx.dt, ctx.saved_units, ctx.x_in = units, dt, saved_units, x_in
        ctx.mirrors = mirrors
        ctx.params, ctx.tap_units = params, tap_units
        return tuple(outputs)

    @staticmethod
    def backward(ctx, *gouts):
        units, dt, params = ctx.units, ctx.dt, ctx.params
        dev = ctx.x_in.device
        sp = _lib.stream_ptr(dev)
        stages = [s for u in units for s in u.stages]
        sizes = [p.numel() for p in params]
        # every parameter gradient of the chain, each on a 32-byte boundary (vector stores / atomics)
        sink = _STATE["grad_sink"]
        sunk = [sink.view_for(p) if sink is not None else None for p in params]
        gflat = torch.zeros(sum((n + 7) // 8 * 8 for n, sv in zip(sizes, sunk) if sv is None), dtype=torch.float32, device=dev)
        gviews, off = [], 0
        for p, n, sv in zip(params, sizes, sunk):
            if sv is not None:
                gviews.append(sv)
                continue
            gviews.append(gflat[off:off + n].view(p.shape))
            off += (n + 7) // 8 * 8
        sum_arena = torch.zeros(sum(2 * s.conv.out_channels for s in stages if s.bn is not None) or 1, dtype=torch.float64, device=dev)
        coef_arena = torch.empty(sum(3 * s.conv.out_channels for s in stages if s.bn is not None) or 1, dtype=torch.float32, device=dev)
        tickets = torch.zeros(len(stages), dtype=torch.int32, device=dev)
        s_off = c_off = k_i = 0
        gout_of = {ui: g for ui, g in zip(ctx.tap_units, gouts) if g is not None}
        need_x_grad = ctx.needs_input_grad[2]

        code = _lib.F32 if dt == torch.float32 else _lib.BF16
        es = 4 if dt == torch.float32 else 2
        p_end = len(params)
        g = None                                         # gradient w.r.t. the current unit's output
        for ui in range(len(units) - 1, -1, -1):
            u = units[ui]
            unit_in, unit_geom, recs = ctx.saved_units[ui]
            gt = gout_of.get(ui)
            if gt is not None:
                gt = _as_nhwc(gt, dt)
                if g is None:
                    g = gt
                else:
                    tmp = torch.empty_like(g)
                    _lib.call("ehgr_row_apply", ctypes.byref(op_plain(g)), gt.data_ptr(), tmp.data_ptr(),
                              g.numel() // g.shape[1], g.shape[1], code, sp, algo_bytes=3 * g.numel() * es)
                    g = tmp
            if g is None:
                raise RuntimeError("chain backward reached a unit without an incoming gradient")
            g_unit_out = g
            p_begin = p_end - sum(st_.n_params() for st_ in u.stages)
            p_off = [p_begin]
            for st_ in u.stages[:-1]:
                p_off.append(p_off[-1] + st_.n_params())
            unit_grad_done = False
            for si in range(len(u.stages) - 1, -1, -1):
                st = u.stages[si]
                raw, vec, geom, tr, act_state, aux = recs[si]
                nt, h, wd, cin = geom
                cout = st.conv.out_channels
                ho, wo = raw.shape[2], raw.shape[3]
                m_out = nt * ho * wo
                pb = p_off[si]
                w, gw = params[pb], gviews[pb]
                if st.bn is None:                        # bare convolution: d(raw) is the incoming gradient itself
                    dy_op = op_plain(g)
                    k_i += 1
                else:
                    gamma = params[pb + 1]
                    ggam, gbet = gviews[pb + 1], gviews[pb + 2]
                sums = sum_arena[s_off:s_off + 2 * cout] if st.bn is not None else None
                if st.bn is not None:
                  s_off += 2 * cout
                  coef = coef_arena[c_off:c_off + 3 * cout].view(3, cout)
                  c_off += 3 * cout
                  bwd = BnBwd(gamma=gamma.data_ptr(), mean=vec[2].data_ptr(), invstd=vec[3].data_ptr(), ca=coef[0].data_ptr(),
                            cb=coef[1].data_ptr(), cc=coef[2].data_ptr(), dgamma=ggam.data_ptr(), dbeta=gbet.data_ptr(),
                            counter=tickets[k_i:].data_ptr(), count=m_out, training=int(tr))
                  k_i += 1
                  _lib.call("ehgr_bn_bwd_reduce_fin", g.data_ptr(), raw.data_ptr(), vec[0].data_ptr(), vec[1].data_ptr(),
                            int(st.relu6), sums.data_ptr(), m_out, cout, code, ctypes.byref(bwd) if FUSED_FINALIZE & 2 else None,
                            sp, algo_bytes=2 * m_out * cout * es)
                  if not FUSED_FINALIZE & 2:
                      _lib.call("ehgr_bn_bwd_finalize", sums.data_ptr(), m_out, gamma.data_ptr(), vec[2].data_ptr(),
                                vec[3].data_ptr(), int(tr), coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(),
                                ggam.data_ptr(), gbet.data_ptr(), cout, sp)
                  dy_op = op_bnbwd(g, raw, coef[0], coef[1], coef[2], vec[0], vec[1], st.relu6)
                if st.kind == 'stem':
                    _lib.call("ehgr_stem_wgrad", ctypes.byref(dy_op), ctx.x_in.data_ptr(), gw.data_ptr(), nt, h, wd,
                              cout, _lib.dtype_code(ctx.x_in), code, sp,
                              algo_bytes=2 * m_out * cout * es + ctx.x_in.numel() * ctx.x_in.element_size(),
                              algo_flops=2 * 27 * m_out * cout)
                    g = None
                    continue
                # forward operand of this stage, re-derived from what was saved
                if st.kind == 'conv3' and aux[2] is not None:        # the materialised upsampled input
                    a_op = op_conv3(aux[2], None, None, 0, ho, wo, cin, False)
                elif st.kind == 'conv3':
                    a_op = (op_conv3(unit_in, None, None, 0, ho, wo, cin, st.up) if si == 0 else
                            op_conv3(recs[si - 1][0], *(recs[si - 1][1][:2] if recs[si - 1][1] is not None else (None, None)),
                                     u.stages[si - 1].relu6, ho, wo, cin, st.up))
                elif act_state is not None:
                    a_op = action_ops.gate_op(act_state)
                elif si == 0:
                    a_op = (op_shift(unit_in, st.shift[0], st.shift[1], h * wd, 1) if st.shift is not None
                            else op_plain(unit_in))
                elif recs[si - 1][1] is None:
                    a_op = op_plain(recs[si - 1][0])
                else:
                    a_op = op_affine(recs[si - 1][0], recs[si - 1][1][0], recs[si - 1][1][1], u.stages[si - 1].relu6)
                m_in = nt * h * wd
                need_dgrad = not (si == 0 and ui == 0 and not need_x_grad)
                g_prev = None
                dy_src = g                               # tensor holding d(raw) once the BN-backward operand is materialised
                if (st.kind == 'pw' and need_dgrad and dy_op.mode != 0) or (st.kind == 'conv3' and dy_op.mode != 0):
                    # dgrad AND wgrad both consume d(raw): evaluate the BN-backward operand once (one
                    # streaming pass at full occupancy) so the tensor-core producers only copy bf16 rows
                    draw = torch.empty_like(raw)
                    _lib.call("ehgr_row_apply", ctypes.byref(dy_op), 0, draw.data_ptr(), m_out, cout, code, sp,
                              algo_bytes=3 * m_out * cout * es)
                    dy_op = op_plain(draw)
                    dy_src = draw
                if st.kind == 'conv3':
                    # wgrad as a [cout, 9*cin] GEMM over the same im2col operand, then back to [cout, cin, 3, 3]
                    dwp = torch.zeros(cout * 9 * cin, dtype=torch.float32, device=dev)
                    _lib.call("ehgr_pw_wgrad", ctypes.byref(dy_op), ctypes.byref(a_op), dwp.data_ptr(), m_out, 9 * cin, cout,
                              code, _STATE["engine"], sp, tag="[conv3x3]",
                              algo_bytes=(m_out * cout + m_in * cin) * es + 36 * cin * cout, algo_flops=2 * m_out * 9 * cin * cout)
                    _lib.call("ehgr_conv3_unpack_grad", dwp.data_ptr(), gw.data_ptr(), cout, cin, sp)
                    if need_dgrad:
                        # dgrad = the same convolution of d(raw) with the flipped, transposed filter
                        wd32, wd16 = aux[1]
                        g_up = _nhwc_empty(nt, ho, wo, cin, dt, dev)
                        _lib.call("ehgr_pw_gemm_w16", ctypes.byref(op_conv3(dy_src, None, None, 0, ho, wo, cout, False)),
                                  wd32.data_ptr(), _lib.ptr(wd16), 0, g_up.data_ptr(), 0, 0, m_out, 9 * cout, cin, code,
                                  _STATE["engine"], sp, tag="[conv3x3]",
                                  algo_bytes=m_out * (cout + cin) * es + 9 * cin * cout * es, algo_flops=2 * m_out * 9 * cin * cout)
                        if st.up:                        # adjoint of the nearest x2 upsample folded into the forward gather
                            g_prev = _nhwc_empty(nt, h, wd, cin, dt, dev)
                            _lib.call("ehgr_upsample2_bwd", g_up.data_ptr(), g_prev.data_ptr(), nt, h, wd, cin, code, sp,
                                      algo_bytes=5 * g_prev.numel() * es)
                        else:
                            g_prev = g_up
                elif st.kind == 'pw':
                    _lib.call("ehgr_pw_wgrad", ctypes.byref(dy_op), ctypes.byref(a_op), gw.data_ptr(), m_in, cin, cout,
                              code, _STATE["engine"], sp,
                              algo_bytes=((1 if dy_op.mode == 0 else 2) * cout + cin) * m_in * es + cin * cout * 4,
                              algo_flops=2 * m_in * cin * cout)
                    if need_dgrad:
                        g_prev = _nhwc_empty(nt, h, wd, cin, dt, dev)
                        # residual units without a shift: fold "+ g_unit_out" into the dgrad epilogue
                        fuse_res = si == 0 and u.residual and st.shift is None and act_state is None
                        w16 = _bf16_mirror(w, dt, ctx.mirrors)
                        _lib.call("ehgr_pw_gemm_w16", ctypes.byref(dy_op), w.data_ptr(), _lib.ptr(w16), 1, g_prev.data_ptr(),
                                  g_unit_out.data_ptr() if fuse_res else 0, 0, m_in, cout, cin, code,
                                  _STATE["engine"], sp,
                                  algo_bytes=((1 if dy_op.mode == 0 else 2) * cout + cin * (2 if fuse_res else 1)) * m_in * es,
                                  algo_flops=2 * m_in * cin * cout)
                else:
                    # fused backward: one shared-memory staging of rowop(dy) and rowop(a) per tile gives
                    # both the input gradient and the weight gradient
                    g_prev = _nhwc_empty(nt, h, wd, cin, dt, dev)
                    _lib.call("ehgr_dw_bwd", ctypes.byref(dy_op), ctypes.byref(a_op), w.data_ptr(), g_prev.data_ptr(),
                              gw.data_ptr(), nt, h, wd, cin, st.stride, code, sp,
                              algo_bytes=(2 * m_out + 2 * m_in) * cin * es, algo_flops=18 * (m_out + m_in) * cin)
                g = g_prev
                if act_state is not None and g is not None:
                    # g is d(loss)/d(gated input): run the ACTION backward (adds the residual gradient)
                    g = action_ops.backward(act_state, params[pb + 3:pb + 13], gviews[pb + 3:pb + 13], g, unit_in,
                                            g_unit_out if u.residual else None)
                    unit_grad_done = True
            # gradient w.r.t. the unit input: undo the shift, add the residual branch
            st0 = u.stages[0]
            if g is not None and st0.kind != 'stem' and not unit_grad_done:
                nt, h, wd, cin = unit_geom
                if st0.shift is not None:
                    gx = torch.empty_like(g)
                    _lib.call("ehgr_row_apply", ctypes.byref(op_shift(g, st0.shift[0], st0.shift[1], h * wd, -1)),
                              g_unit_out.data_ptr() if u.residual else 0, gx.data_ptr(), nt * h * wd, cin, code, sp,
                              algo_bytes=(2 + int(u.residual)) * g.numel() * es)
                    g = gx
                elif u.residual and st0.kind != 'pw':
                    gx = torch.empty_like(g)
                    _lib.call("ehgr_row_apply", ctypes.byref(op_plain(g)), g_unit_out.data_ptr(), gx.data_ptr(),
                              nt * h * wd, cin, code, sp, algo_bytes=3 * g.numel() * es)
                    g = gx
            if sink is not None:   # this unit's parameter gradients are complete (stream order): buckets may go
                sink.mark_done([p for p, sv in zip(params[p_begin:p_end], sunk[p_begin:p_end]) if sv is not None])
            p_end = p_begin
        gx = None
        if need_x_grad and g is not None:
            gx = g if g.dtype == ctx.x_in.dtype else g.to(ctx.x_in.dtype)
        pg = [gv if (p.requires_grad and sv is None) else None for gv, p, sv in zip(gviews, params, sunk)]
        return (None, None, gx, *pg)


def _chain_params(units):
    ps = []
    for u in units:
        for s in u.stages:
            ps += [s.conv.weight] + ([s.bn.weight, s.bn.bias] if s.bn is not None else [])
            if s.action is not None:
                ps += action_ops.action_params(s.action)
    return ps


def run_chain(units: List[Unit], x):
    return _ChainFunction.apply(units, _STATE["dtype"], x, *_chain_params(units))


# ------------------------------------------------------------------------------------------------
# module entry points
# ------------------------------------------------------------------------------------------------
def inverted_residual(m, x):
    """InvertedResidual.forward (archs/mobilenet_v2.py:62-66) as a one-unit chain."""
    _lib.require_cuda(x)
    return run_chain([unit_of_block(m)], x)[0]


def sepconv_stack(modules, x):
    """A stack of SepConv blocks (one exit head's ``scala``: models/models_SD.py:214-253) as ONE fused chain."""
    _lib.require_cuda(x)
    return run_chain([unit_of_sepconv(m) for m in modules], x)[0]


def mobilenet_v2_features(model, x, taps: Sequence[int] = ()):
    """features[0..18] -> [NT, 1280, H/32, W/32] (NHWC strides).  With ``taps`` returns a tuple: the
    outputs of the listed feature indices (ascending) followed by the final map."""
    _lib.require_cuda(x)
    outs = run_chain(units_of_backbone(model, taps), x)
    return outs[0] if not taps else outs


def mobilenet_v2_forward(model, x):
    f = mobilenet_v2_features(model, x)
    return model.classifier(global_avg_pool(f))


class _PoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        nt, c, h, w = x.shape
        ctx.shape, ctx.dtype = (nt, c, h, w), x.dtype
        pooled = torch.empty((nt, c), dtype=torch.float32, device=x.device)
        _lib.call("ehgr_pool_fwd", ctypes.byref(op_plain(x)), pooled.data_ptr(), nt, h * w, c, _lib.dtype_code(x),
                  _lib.stream_ptr(x.device), algo_bytes=x.numel() * x.element_size())
        return pooled

    @staticmethod
    def backward(ctx, g):
        nt, c, h, w = ctx.shape
        da = _nhwc_empty(nt, h, w, c, ctx.dtype, g.device)
        g = g.contiguous().float()
        _lib.call("ehgr_pool_bwd", g.data_ptr(), da.data_ptr(), nt, h * w, c, _lib.dtype_code(da),
                  _lib.stream_ptr(g.device), algo_bytes=da.numel() * da.element_size())
        return da


def global_avg_pool(x):
    """x.mean(3).mean(2) (archs/mobilenet_v2.py:112) -> fp32 [NT, C]."""
    _lib.require_cuda(x)
    dt = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
    return _PoolFunction.apply(_as_nhwc(x, dt))


class _FcConsensusFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, weight, bias, n_segment):
        nt, f = feat.shape
        n, k = nt // n_segment, weight.shape[0]
        feat = feat.contiguous().float()
        w = weight.contiguous().float()
        b = bias.contiguous().float() if bias is not None else None
        meanfeat = torch.empty((n, f), dtype=torch.float32, device=feat.device)
        logits = torch.empty((n, k), dtype=torch.float32, device=feat.device)
        _lib.call("ehgr_fc_consensus_fwd", feat.data_ptr(), w.data_ptr(), _lib.ptr(b), meanfeat.data_ptr(),
                  logits.data_ptr(), n, n_segment, f, k, _lib.stream_ptr(feat.device))
        ctx.save_for_backward(meanfeat, w)
        ctx.meta = (n, n_segment, f, k, bias is not None)
        return logits

    @staticmethod
    def backward(ctx, g):
        meanfeat, w = ctx.saved_tensors
        n, T, f, k, has_bias = ctx.meta
        g = g.contiguous().float()
        dfeat = torch.empty((n * T, f), dtype=torch.float32, device=g.device)
        dw = torch.zeros_like(w)
        db = torch.zeros(k, dtype=torch.float32, device=g.device) if has_bias else None
        _lib.call("ehgr_fc_consensus_bwd", g.data_ptr(), meanfeat.data_ptr(), w.data_ptr(), dfeat.data_ptr(),
                  dw.data_ptr(), _lib.ptr(db), n, T, f, k, _lib.stream_ptr(g.device))
        return dfeat, dw, db, None


def fc_consensus(feat, linear: nn.Linear, n_segment: int):
    """new_fc followed by the 'avg' segment consensus: [N*T, F] -> [N, K]."""
    if feat.shape[0] % n_segment:
        raise RuntimeError(f"shape '[-1, {n_segment}, ...]' is invalid for input with {feat.shape[0]} rows")
    return _FcConsensusFunction.apply(feat, linear.weight, linear.bias, n_segment)


def classifier_head(tsn, fmap):
    """global average pool -> Dropout -> new_fc -> mean over segments (models/models.py:341-356)."""
    pooled = global_avg_pool(fmap)
    drop = getattr(tsn.base_model, tsn.base_model.last_layer_name)
    if isinstance(drop, nn.Dropout):
        pooled = drop(pooled)                       # mask-and-scale on the tiny [NT, F] tensor
        if tsn.consensus_type == 'avg' and tsn.new_fc is not None and tsn.before_softmax:
            return fc_consensus(pooled, tsn.new_fc, tsn.num_segments)
        z = tsn.new_fc(pooled)
    else:                                           # dropout == 0: the Linear sits in the backbone
        z = drop(pooled)
    if not tsn.before_softmax:
        z = tsn.softmax(z)
    z = z.view((-1, tsn.num_segments) + z.size()[1:])
    return tsn.consensus(z).squeeze(1)


# ------------------------------------------------------------------------------------------------
# MTMM depth decoder (models/models_MTMM.py:129-155) on the fused chain
# ------------------------------------------------------------------------------------------------
def parse_depth_decoder(seq: nn.Sequential):
    """[Conv3x3, BN, ReLU, (Upsample x2)]* , Conv1x1(bias), Sigmoid  ->  (conv3 stages, head conv) or None when the module
    does not have that shape (then the caller runs it as the nn.Sequential it is)."""
    mods = list(seq)
    stages, i, up = [], 0, False
    while i + 2 < len(mods):
        conv, bn, act = mods[i], mods[i + 1], mods[i + 2]
        if not (isinstance(conv, nn.Conv2d) and conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1)
                and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None and isinstance(bn, nn.BatchNorm2d)
                and isinstance(act, nn.ReLU) and conv.in_channels % 8 == 0 and conv.out_channels % 8 == 0):
            break
        st = _conv_bn_stage('conv3', conv, bn, 2)
        st.up = up
        stages.append(st)
        i += 3
        up = False
        if i < len(mods) and isinstance(mods[i], nn.Upsample):
            if mods[i].mode != 'nearest' or float(mods[i].scale_factor) != 2.0:
                return None
            up = True
            i += 1
    if up or not stages or len(mods) != i + 2:
        return None
    head, sig = mods[i], mods[i + 1]
    if not (isinstance(head, nn.Conv2d) and head.kernel_size == (1, 1) and head.out_channels == 1 and head.groups == 1
            and isinstance(sig, nn.Sigmoid) and head.in_channels % 8 == 0 and head.in_channels <= 256):
        return None
    return stages, head


class _DepthHeadFunction(torch.autograd.Function):
    """sigmoid(Conv2d(C, 1, 1, bias)(x)) — csrc/conv3.cu depth_head_kernel; x NHWC in the compute dtype, out fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        nt, c, h, w = x.shape
        out = torch.empty((nt, 1, h, w), dtype=torch.float32, device=x.device)
        wv = weight.detach().reshape(-1).float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        _lib.call("ehgr_depth_head_fwd", ctypes.byref(op_plain(x)), wv.data_ptr(), _lib.ptr(b), out.data_ptr(), nt * h * w, c,
                  _lib.dtype_code(x), _lib.stream_ptr(x.device), algo_bytes=x.numel() * x.element_size() + out.numel() * 4)
        ctx.save_for_backward(x, wv, out)
        ctx.has_bias, ctx.wshape = bias is not None, weight.shape
        return out

    @staticmethod
    def backward(ctx, g):
        x, wv, out = ctx.saved_tensors
        nt, c, h, w = x.shape
        g = g.contiguous().float()
        gx = torch.empty_like(x)
        dw = torch.zeros(c, dtype=torch.float32, device=x.device)
        db = torch.zeros(1, dtype=torch.float32, device=x.device) if ctx.has_bias else None
        _lib.call("ehgr_depth_head_bwd", ctypes.byref(op_plain(x)), wv.data_ptr(), out.data_ptr(), g.data_ptr(), gx.data_ptr(),
                  dw.data_ptr(), _lib.ptr(db), nt * h * w, c, _lib.dtype_code(x), _lib.stream_ptr(x.device),
                  algo_bytes=2 * x.numel() * x.element_size() + 2 * out.numel() * 4)
        return gx, dw.view(ctx.wshape), db


def depth_decoder(seq: nn.Sequential, fmap):
    """``global_decoder(fmap)`` of models/models_MTMM.py:129-155 -> [NT, 1, 8H, 8W] fp32: the four 3x3 convolutions are ONE
    fused chain of implicit-GEMM units (tcgen05, im2col by TMA boxes; BatchNorm statistics in the GEMM epilogue,
    BatchNorm+ReLU applied once per unit, the three nearest x2 upsamples as one small copy each), the 1x1 convolution +
    sigmoid is one more kernel.  No library convolution / BatchNorm / activation / upsample op."""
    _lib.require_cuda(fmap)
    parsed = parse_depth_decoder(seq)
    if parsed is None:
        return seq(fmap)                                 # another decoder architecture (e.g. ConvTranspose2d): library modules
    stages, head = parsed
    # one unit per convolution: a unit materialises BatchNorm+ReLU once (a few MB), so the nine taps of the next
    # convolution gather finished activations instead of re-evaluating the affine map nine times per pixel
    y = run_chain([Unit([st]) for st in stages], fmap)[0]
    return _DepthHeadFunction.apply(y, head.weight, head.bias)


# ------------------------------------------------------------------------------------------------
# stand-alone ACTION module (inside an InvertedResidual the chain fuses the gating into the GEMM A-load)
# ------------------------------------------------------------------------------------------------
def action_forward(m, x):
    """out = net(x_shift * (3 + g_STE + g_CE + g_ME)) — reference models/action.py:61-116.  The gated tensor
    comes from the fused kernels (csrc/action.cu); ``m.net`` is whatever module was wrapped."""
    _lib.require_cuda(x)
    dt = _STATE["dtype"]
    y = action_ops.gated(m, x, dt)
    if y.dtype != x.dtype and not torch.is_autocast_enabled():
        y = y.to(x.dtype)
    return m.net(y)
