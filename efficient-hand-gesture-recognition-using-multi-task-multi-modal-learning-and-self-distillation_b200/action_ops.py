"""Host side of the fused ACTION kernels (csrc/action.cu): buffer management and the call sequences.

forward : xs, small reductions            (ehgr_action_xs)
          BatchNorm of the motion squeeze (ehgr_bn_finalize on q's statistics)
          gates g1 / g2 / g3              (ehgr_action_gates)
          y = xs * (3 + g1 + g2 + g3)     — inside the wrapped 1x1 conv's GEMM as a GATE row operand (chain),
                                            or materialised by ehgr_row_apply (stand-alone ``Action``)
backward: gy = d(loss)/dy  ->  reductions -> per-clip gate backward -> BN backward -> d(xs) -> FIR adjoint

Reference: models/action.py:61-116 (forward; backward is autograd's there).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import ActionArgs, RowOp

# order of the module's parameters as they are passed through autograd Functions
PARAM_NAMES = ("action_shift.weight", "action_p1_conv1.weight", "action_p2_squeeze.weight", "action_p2_conv1.weight",
               "action_p2_expand.weight", "action_p3_squeeze.weight", "action_p3_bn1.weight", "action_p3_bn1.bias",
               "action_p3_conv1.weight", "action_p3_expand.weight")


def action_params(mod):
    return [mod.action_shift.weight, mod.action_p1_conv1.weight, mod.action_p2_squeeze.weight, mod.action_p2_conv1.weight,
            mod.action_p2_expand.weight, mod.action_p3_squeeze.weight, mod.action_p3_bn1.weight, mod.action_p3_bn1.bias,
            mod.action_p3_conv1.weight, mod.action_p3_expand.weight]


def _pad(n_: int) -> int:
    """Sub-buffers of one allocation start on 32-byte boundaries (vector loads of the per-channel gates)."""
    return (n_ + 7) // 8 * 8


class ActionState:
    """Everything the backward needs from one forward call."""
    __slots__ = ("n", "t", "h", "w", "c", "cr", "xs", "small", "views", "bn3_vec", "bn3_training", "dtype")


def _args(st: ActionState, params, **extra) -> ActionArgs:
    a = ActionArgs(n=st.n, t=st.t, h=st.h, w=st.w, c=st.c, cr=st.cr)
    sw, p1, sq2, c1, ex2, sq3, _g3, _b3, c3, ex3 = params
    a.shift_w, a.p1_w, a.p2_squeeze, a.p2_conv1, a.p2_expand = (sw.data_ptr(), p1.data_ptr(), sq2.data_ptr(),
                                                                c1.data_ptr(), ex2.data_ptr())
    a.p3_squeeze, a.p3_conv1, a.p3_expand = sq3.data_ptr(), c3.data_ptr(), ex3.data_ptr()
    for k, v in st.views.items():
        setattr(a, k, v.data_ptr())
    a.bn3_scale, a.bn3_shift = st.bn3_vec[0].data_ptr(), st.bn3_vec[1].data_ptr()
    for k, v in extra.items():
        setattr(a, k, v.data_ptr())
    return a


def forward_gates(mod, x, params, dt):
    """x: NHWC-strided [NT,C,H,W] tensor of dtype dt.  Returns (ActionState, GATE RowOp)."""
    nt, c, h, w = x.shape
    T = mod.n_segment
    if nt % T:
        raise RuntimeError(f"shape '[{nt // T}, {T}, {c}, {h}, {w}]' is invalid for input of size {x.numel()}")
    st = ActionState()
    st.n, st.t, st.h, st.w, st.c, st.cr, st.dtype = nt // T, T, h, w, c, mod.reduced_channels, dt
    if st.cr < 1:
        raise RuntimeError("Action needs at least 16 input channels (reduced_channels = C // 16)")
    dev = x.device
    M, cr = nt * h * w, st.cr
    sizes = {"mrow": M, "pool": nt * c, "q": M * cr, "g1": M, "g2": nt * c, "g3": nt * c, "s": nt * cr, "u": nt * cr,
             "pi": nt * cr}
    st.small = torch.empty(sum(_pad(v) for v in sizes.values()), dtype=torch.float32, device=dev)
    st.views, off = {}, 0
    for k, n_ in sizes.items():
        st.views[k] = st.small[off:off + n_]
        off += _pad(n_)
    qstats = torch.zeros(2 * cr, dtype=torch.float64, device=dev)
    st.views["qstats"] = qstats
    st.bn3_vec = torch.empty((4, cr), dtype=torch.float32, device=dev)
    st.xs = torch.empty_like(x)
    bn3 = mod.action_p3_bn1
    st.bn3_training = bn3.training
    sp = _lib.stream_ptr(dev)
    code = _lib.dtype_code(x)
    es = x.element_size()
    a = _args(st, params)
    _lib.call("ehgr_action_xs", ctypes.byref(a), x.data_ptr(), st.xs.data_ptr(), code, sp, algo_bytes=2 * x.numel() * es)
    _lib.call("ehgr_bn_finalize", qstats.data_ptr(), M, params[6].data_ptr(), params[7].data_ptr(),
              bn3.running_mean.data_ptr(), bn3.running_var.data_ptr(), float(bn3.momentum), float(bn3.eps),
              int(bn3.training), st.bn3_vec[0].data_ptr(), st.bn3_vec[1].data_ptr(), st.bn3_vec[2].data_ptr(),
              st.bn3_vec[3].data_ptr(), cr, sp)
    if bn3.training and bn3.num_batches_tracked is not None:
        bn3.num_batches_tracked.add_(1)
    _lib.call("ehgr_action_gates", ctypes.byref(a), sp)
    return st, gate_op(st)


def gate_op(st: ActionState) -> RowOp:
    v = st.views
    return RowOp(mode=4, in1=st.xs.data_ptr(), in2=v["g1"].data_ptr(), scale=v["g2"].data_ptr(), shift=v["g3"].data_ptr(),
                 hw=st.h * st.w)


def backward(st: ActionState, params, grads, gy, x, addend):
    """gy: gradient w.r.t. the gated tensor y (NHWC, dtype st.dtype); grads: ten fp32 tensors shaped like
    `params`, zero-initialised, receiving the parameter gradients; returns dx (+ addend)."""
    dev = gy.device
    nt, c, cr = st.n * st.t, st.c, st.cr
    M = nt * st.h * st.w
    ws = torch.empty(2 * _pad(M) + 2 * _pad(nt * c) + _pad(nt * cr), dtype=torch.float32, device=dev)
    o = [0]

    def take(n_):
        v = ws[o[0]:o[0] + n_]
        o[0] += _pad(n_)
        return v
    extra = {"dg1": take(M), "dm": take(M), "dgc": take(nt * c), "dpool": take(nt * c), "dd": take(nt * cr)}
    sums = torch.zeros(2 * cr, dtype=torch.float64, device=dev)
    coef = torch.empty((3, cr), dtype=torch.float32, device=dev)
    extra["bn3_sums"] = sums
    extra.update(bn3_ca=coef[0], bn3_cb=coef[1], bn3_cc=coef[2])
    gsw, gp1, gsq2, gc1, gex2, gsq3, gg3, gb3, gc3, gex3 = grads
    extra.update(d_shift_w=gsw, d_p1_w=gp1, d_p2_squeeze=gsq2, d_p2_conv1=gc1, d_p2_expand=gex2, d_p3_squeeze=gsq3,
                 d_p3_conv1=gc3, d_p3_expand=gex3)
    a = _args(st, params, **extra)
    sp = _lib.stream_ptr(dev)
    code = _lib.dtype_code(gy)
    es = gy.element_size()
    _lib.call("ehgr_action_bwd_reduce", ctypes.byref(a), gy.data_ptr(), st.xs.data_ptr(), code, sp,
              algo_bytes=2 * gy.numel() * es)
    _lib.call("ehgr_action_bwd_small", ctypes.byref(a), sp)
    _lib.call("ehgr_bn_bwd_finalize", sums.data_ptr(), M, params[6].data_ptr(), st.bn3_vec[2].data_ptr(),
              st.bn3_vec[3].data_ptr(), int(st.bn3_training), coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(),
              gg3.data_ptr(), gb3.data_ptr(), cr, sp)
    dxs = torch.empty_like(gy)
    _lib.call("ehgr_action_bwd_dxs", ctypes.byref(a), gy.data_ptr(), st.xs.data_ptr(), dxs.data_ptr(), code, sp,
              algo_bytes=3 * gy.numel() * es)
    dx = torch.empty_like(gy)
    _lib.call("ehgr_action_fir_bwd", ctypes.byref(a), dxs.data_ptr(), x.data_ptr(), _lib.ptr(addend), dx.data_ptr(), code, sp,
              algo_bytes=(3 + int(addend is not None)) * gy.numel() * es)
    return dx


class _ActionGateFunction(torch.autograd.Function):
    """Stand-alone ``Action``: y = xs * (3 + g1 + g2 + g3) materialised (the wrapped ``net`` is arbitrary)."""

    @staticmethod
    def forward(ctx, mod, dt, x, *params):
        from . import fused
        xin = fused._as_nhwc(x, dt)
        p32 = [p.detach().contiguous().float() for p in params]
        st, op = forward_gates(mod, xin, p32, dt)
        y = torch.empty_like(xin)
        nt, c, h, w = xin.shape
        _lib.call("ehgr_row_apply", ctypes.byref(op), 0, y.data_ptr(), nt * h * w, c, _lib.dtype_code(y),
                  _lib.stream_ptr(y.device), algo_bytes=2 * y.numel() * y.element_size())
        ctx.st, ctx.x, ctx.p32, ctx.params, ctx.in_dtype = st, xin, p32, params, x.dtype
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import fused
        gy = fused._as_nhwc(gy, ctx.st.dtype)
        grads = [torch.zeros_like(p) for p in ctx.p32]
        dx = backward(ctx.st, ctx.p32, grads, gy, ctx.x, None)
        pg = [g.view(p.shape).to(p.dtype) if p.requires_grad else None for g, p in zip(grads, ctx.params)]
        return (None, None, dx.to(ctx.in_dtype), *pg)


def gated(mod, x, dt):
    """The input of ``mod.net``: x_p1 + x_p2 + x_p3 of the reference (models/action.py:115)."""
    return _ActionGateFunction.apply(mod, dt, x, *action_params(mod))
