"""N3 — torchvision ResNet bottleneck backbones (ResNet-50 / -101 / -152) on the library's own sm_100a kernels.

The reference builds ``torchvision.models.resnet50`` for ``base_model='resnet50'`` (models/models.py:108-117) and
``make_temporal_shift`` wraps ``conv1`` of every ``Bottleneck`` (models/temporal_shift.py:101-146, place='blockres').
The torchvision module tree stays the parameter container (names = checkpoint contract); one
``torch.autograd.Function`` runs stem -> max-pool -> layer1..4, forward and backward, through the C ABI:

  * ``conv1`` 7x7/2 + bn1 statistics            bf16: ehgr_stem7_im2col + ehgr_pw_gemm_bn (tcgen05); fp32: ehgr_stem7_fwd
  * ReLU + MaxPool2d(3, 2, 1)                   ehgr_maxpool3_fwd/bwd     (lazy BatchNorm+ReLU applied on load)
  * Bottleneck.conv1 (1x1, TemporalShift)       ehgr_pw_gemm_bn, SHIFT / PLAIN row operand (tcgen05 for bf16)
  * Bottleneck.conv2 (3x3)                      ehgr_pw_gemm_bn, CONV3 row operand (implicit GEMM, im2col by TMA boxes);
                                                stride 2 = the stride-1 convolution sampled at even pixels
                                                (ehgr_subsample2_fwd, which also takes the BatchNorm statistics)
  * Bottleneck.conv3 (1x1)                      ehgr_pw_gemm_bn, AFFINE operand (bn2 + ReLU applied on load)
  * downsample (1x1 stride s + BN)              ehgr_subsample2_fwd + ehgr_pw_gemm_bn
  * relu(bn3(.) + identity)                     ehgr_bn_add_relu / ehgr_relu_bwd
  * backward: BatchNorm-backward reductions, dgrad / wgrad GEMMs as in fused._ChainFunction (csrc/bn.cu, pw_tc*.cu)

Not covered (``supported()`` says why; TSN.forward then runs the torchvision modules and warns): BasicBlock nets,
grouped / dilated bottlenecks, ``Action`` on widths above 256 channels (csrc/action.cu keeps Cr <= 16),
place='block', temporal_pool.  The input gradient (d loss / d video) is not computed.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib, fused
from .fused import (_bf16_mirror, _nhwc_empty, _pack_conv3, op_affine, op_bnbwd, op_conv3, op_plain, op_shift)


# The 7x7 stem as a tensor-core GEMM over a materialised patch matrix (ehgr_stem7_im2col, K = 147 padded to 160) instead of
# the exact-fp32 CUDA-core kernel.  None = by storage dtype (bf16 -> GEMM, fp32 -> CUDA cores: the parity mode).  Measured
# on B200 at 1024 frames of 224^2: the CUDA-core pair cost 28 + 25 ms of a 168 ms step (profiles/README.md, round 2 N3).
STEM_GEMM = None
STEM_KP = 160


def _stem_as_gemm(dt) -> bool:
    return (dt == torch.bfloat16 and fused._STATE["engine"] != 1) if STEM_GEMM is None else bool(STEM_GEMM)


@dataclass
class _Block:
    conv1: nn.Conv2d
    bn1: nn.BatchNorm2d
    conv2: nn.Conv2d
    bn2: nn.BatchNorm2d
    conv3: nn.Conv2d
    bn3: nn.BatchNorm2d
    down: Optional[Tuple[nn.Conv2d, nn.BatchNorm2d]]
    stride: int
    shift: Optional[Tuple[int, int]]     # (n_segment, fold) of the TemporalShift around conv1
    stage: int                           # 1..4
    last_of_stage: bool = False

    def modules(self):
        ms = [self.conv1, self.bn1, self.conv2, self.bn2, self.conv3, self.bn3]
        return ms + (list(self.down) if self.down is not None else [])


@dataclass
class _Plan:
    conv1: nn.Conv2d
    bn1: nn.BatchNorm2d
    blocks: List[_Block]


def _default_bn(bn) -> bool:
    return isinstance(bn, nn.BatchNorm2d) and bn.affine and bn.track_running_stats and bn.momentum is not None


def _plain_conv(c, k, stride_ok=(1,)) -> bool:
    return (isinstance(c, nn.Conv2d) and c.kernel_size == (k, k) and c.stride[0] == c.stride[1] and c.stride[0] in stride_ok
            and c.padding == (k // 2, k // 2) and c.dilation == (1, 1) and c.groups == 1 and c.bias is None)


def plan_of(model) -> _Plan:
    """torchvision ResNet (Bottleneck) -> execution plan; raises NotImplementedError with the reason when the module tree is
    not one the kernels cover."""
    import torchvision
    from torchvision.models.resnet import Bottleneck
    from .temporal_shift import TemporalShift
    if not isinstance(model, torchvision.models.ResNet):
        raise NotImplementedError(f"not a torchvision ResNet: {type(model).__name__}")
    if not (_plain_conv(model.conv1, 7, (2,)) and model.conv1.in_channels == 3 and model.conv1.out_channels == 64
            and _default_bn(model.bn1)):
        raise NotImplementedError("stem is not Conv2d(3, 64, 7, stride=2, padding=3, bias=False) + BatchNorm2d")
    mp = model.maxpool
    if not (isinstance(mp, nn.MaxPool2d) and mp.kernel_size in (3, (3, 3)) and mp.stride in (2, (2, 2)) and mp.padding in (1, (1, 1))
            and mp.dilation in (1, (1, 1)) and not mp.ceil_mode):
        raise NotImplementedError("max-pool is not MaxPool2d(3, stride=2, padding=1)")
    blocks: List[_Block] = []
    for si, name in enumerate(("layer1", "layer2", "layer3", "layer4"), 1):
        stage = getattr(model, name)
        if not isinstance(stage, nn.Sequential):
            raise NotImplementedError(f"{name} is wrapped by {type(stage).__name__} (temporal_pool is not on this path)")
        for b in stage:
            if not isinstance(b, Bottleneck):
                raise NotImplementedError(f"{name} holds {type(b).__name__}; only torchvision Bottleneck blocks are covered")
            c1, shift = b.conv1, None
            if isinstance(c1, TemporalShift):
                shift, c1 = (c1.n_segment, c1.net.in_channels // c1.fold_div), c1.net
            if not isinstance(c1, nn.Conv2d):
                raise NotImplementedError(f"{name}: conv1 is wrapped by {type(c1).__name__} (Action covers widths up to 256 "
                                          "channels only: csrc/action.cu)")
            if not (_plain_conv(c1, 1) and _plain_conv(b.conv2, 3, (1, 2)) and _plain_conv(b.conv3, 1)
                    and all(_default_bn(x) for x in (b.bn1, b.bn2, b.bn3))):
                raise NotImplementedError(f"{name}: grouped / dilated / biased bottleneck convolutions are not covered")
            down = None
            if b.downsample is not None:
                d = b.downsample
                if not (isinstance(d, nn.Sequential) and len(d) == 2 and _plain_conv(d[0], 1, (1, 2)) and _default_bn(d[1])
                        and d[0].stride[0] == b.conv2.stride[0]):
                    raise NotImplementedError(f"{name}: downsample is not Conv2d(1x1, stride) + BatchNorm2d")
                down = (d[0], d[1])
            elif b.conv2.stride[0] != 1 or c1.in_channels != b.conv3.out_channels:
                raise NotImplementedError(f"{name}: identity shortcut with a shape change")
            if any(c.in_channels % 8 or c.out_channels % 8 for c in (c1, b.conv2, b.conv3)):
                raise NotImplementedError("channel counts must be multiples of 8")
            blocks.append(_Block(c1, b.bn1, b.conv2, b.bn2, b.conv3, b.bn3, down, b.conv2.stride[0], shift, si))
        blocks[-1].last_of_stage = True
    return _Plan(model.conv1, model.bn1, blocks)


def supported(model) -> Tuple[bool, str]:
    try:
        plan_of(model)
        return True, ""
    except NotImplementedError as e:
        return False, str(e)


def _plan_params(plan: _Plan):
    ps = [plan.conv1.weight, plan.bn1.weight, plan.bn1.bias]
    for b in plan.blocks:
        for m in b.modules():
            ps += [m.weight] if isinstance(m, nn.Conv2d) else [m.weight, m.bias]
    return ps


def _half(v: int) -> int:
    return (v - 1) // 2 + 1


class _Arena:
    """Bump allocator over one zero-initialised (float64) or uninitialised (float32) buffer: the per-layer statistics and
    coefficient vectors of a pass are slices of it (one allocation / memset instead of one per layer)."""

    def __init__(self, n, dtype, device, zero):
        self.buf = (torch.zeros if zero else torch.empty)(max(int(n), 1), dtype=dtype, device=device)
        self.off = 0

    def take(self, n):
        out = self.buf[self.off:self.off + n]
        self.off += n
        return out


class _ResNetFunction(torch.autograd.Function):
    """forward(plan, dt, taps, x, *params) -> (outputs of the tapped stages..., layer4 output), NHWC strides."""

    @staticmethod
    def forward(ctx, plan: _Plan, dt, taps: Tuple[int, ...], x, *params):
        ctx.set_materialize_grads(False)
        _lib.require_cuda(x)
        dev = x.device
        sp = _lib.stream_ptr(dev)
        eng = fused._STATE["engine"]
        code = _lib.F32 if dt == torch.float32 else _lib.BF16
        es = 4 if dt == torch.float32 else 2
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"ResNet stem expects [NT,3,H,W], got {tuple(x.shape)}")
        x_in = x.detach().contiguous()
        if x_in.dtype not in (torch.float32, torch.bfloat16):
            x_in = x_in.float()
        nt, _, H, W = x_in.shape
        bns = [plan.bn1] + [m for b in plan.blocks for m in b.modules() if isinstance(m, nn.BatchNorm2d)]
        stat_arena = _Arena(sum(2 * bn.num_features for bn in bns if bn.training), torch.float64, dev, True)
        vec_arena = _Arena(sum(4 * bn.num_features for bn in bns), torch.float32, dev, False)
        mirrors, nbt = {}, []

        def bn_forward(launch, bn, gamma, beta, count):
            """`launch(stats_ptr)` enqueues the producing kernel (it accumulates the batch statistics); returns
            (vec[4, C] = scale, shift, mean, invstd ; training flag)."""
            c = bn.num_features
            tr = bn.training
            stats = stat_arena.take(2 * c) if tr else None
            launch(_lib.ptr(stats))
            vec = vec_arena.take(4 * c).view(4, c)
            _lib.call("ehgr_bn_finalize", _lib.ptr(stats), count, gamma.data_ptr(), beta.data_ptr(), bn.running_mean.data_ptr(),
                      bn.running_var.data_ptr(), float(bn.momentum), float(bn.eps), int(tr), vec[0].data_ptr(), vec[1].data_ptr(),
                      vec[2].data_ptr(), vec[3].data_ptr(), c, sp)
            if tr and bn.num_batches_tracked is not None:
                nbt.append(bn.num_batches_tracked)
            return vec, tr

        def gemm(a_op, w32, w16, out, stats_p, m, k, n, tag="", algo_in=0):
            _lib.call("ehgr_pw_gemm_bn", ctypes.byref(a_op), w32.data_ptr(), _lib.ptr(w16), 0, out.data_ptr(), 0, stats_p, m, k, n,
                      code, eng, None, sp, tag=tag, algo_bytes=(algo_in + m * n) * es + k * n * es, algo_flops=2 * m * k * n)

        # ---- stem: conv1 7x7/2 -> bn1 -> ReLU -> MaxPool2d(3, 2, 1)
        w0, g0, b0 = params[0:3]
        ho, wo = _half(H), _half(W)
        raw0 = _nhwc_empty(nt, ho, wo, 64, dt, dev)
        if _stem_as_gemm(dt):
            # patch matrix [M, 160] (147 taps + zero pad) -> one plain GEMM on the tensor cores; backward's weight-gradient
            # GEMM reads the same matrix, so it is kept (M * 320 bytes = 4 MB per 224^2 frame, ~10 % of what the step saves)
            m0 = nt * ho * wo
            patches = torch.empty((m0, STEM_KP), dtype=dt, device=dev)
            _lib.call("ehgr_stem7_im2col", x_in.data_ptr(), patches.data_ptr(), nt, H, W, STEM_KP, _lib.dtype_code(x_in), code, sp,
                      algo_bytes=x_in.numel() * x_in.element_size() + patches.numel() * es)
            wp32 = torch.empty((64, STEM_KP), dtype=torch.float32, device=dev)
            _lib.call("ehgr_stem7_pack", w0.data_ptr(), wp32.data_ptr(), 64, STEM_KP, _lib.F32, sp)
            wp16 = None
            if dt == torch.bfloat16 and eng != 1:
                wp16 = torch.empty((64, STEM_KP), dtype=torch.bfloat16, device=dev)
                _lib.call("ehgr_stem7_pack", w0.data_ptr(), wp16.data_ptr(), 64, STEM_KP, _lib.BF16, sp)
            vec0, tr0 = bn_forward(lambda st: gemm(op_plain(patches), wp32, wp16, raw0, st, m0, STEM_KP, 64, "[stem7x7]",
                                                   m0 * STEM_KP), plan.bn1, g0, b0, m0)
        else:
            patches = None
            vec0, tr0 = bn_forward(lambda st: _lib.call(
                "ehgr_stem7_fwd", x_in.data_ptr(), w0.data_ptr(), raw0.data_ptr(), st, nt, H, W, 64, _lib.dtype_code(x_in), code, sp,
                algo_bytes=x_in.numel() * x_in.element_size() + raw0.numel() * es, algo_flops=2 * 147 * raw0.numel()),
                plan.bn1, g0, b0, nt * ho * wo)
        h, w = _half(ho), _half(wo)
        cur = _nhwc_empty(nt, h, w, 64, dt, dev)
        pool_idx = torch.empty((nt, h, w, 64), dtype=torch.uint8, device=dev)
        _lib.call("ehgr_maxpool3_fwd", ctypes.byref(op_affine(raw0, vec0[0], vec0[1], 2)), cur.data_ptr(), pool_idx.data_ptr(), nt,
                  ho, wo, 64, code, sp, algo_bytes=raw0.numel() * es + cur.numel() * (es + 1))
        stem_saved = (raw0, vec0, tr0, pool_idx, (ho, wo), patches)

        saved, outputs, tap_blocks = [], [], []
        if 0 in taps:                     # the max-pool output (models/models_MTMM_SD.py:431-476 feeds it to local_decoder)
            outputs.append(cur)
            tap_blocks.append(-1)
            cur = cur.detach()            # layer1.0 saves an alias, not the Function output
        p_i = 3
        for bi, b in enumerate(plan.blocks):
            w1, g1, b1, w2, g2, b2, w3, g3, b3 = params[p_i:p_i + 9]
            p_i += 9
            wd = gd = bd = None
            if b.down is not None:
                wd, gd, bd = params[p_i:p_i + 3]
                p_i += 3
            cin, p, cout = b.conv1.in_channels, b.conv1.out_channels, b.conv3.out_channels
            if cur.shape[1] != cin:
                raise RuntimeError(f"bottleneck expects {cin} input channels, got {cur.shape[1]}")
            m = nt * h * w
            x_blk = cur
            # conv1 (1x1) — the temporal shift is the SHIFT row operand of the GEMM (never materialised)
            a_op = op_shift(x_blk, b.shift[0], b.shift[1], h * w, 1) if b.shift is not None else op_plain(x_blk)
            if b.shift is not None and nt % b.shift[0]:
                raise RuntimeError(f"shape '[-1, {b.shift[0]}, ...]' is invalid for input with {nt} frames")
            raw1 = _nhwc_empty(nt, h, w, p, dt, dev)
            vec1, tr1 = bn_forward(lambda st: gemm(a_op, w1, _bf16_mirror(w1, dt, mirrors), raw1, st, m, cin, p, algo_in=m * cin),
                                   b.bn1, g1, b1, m)
            # bn1 + ReLU once (finished activation): the nine taps of conv2 then gather it with TMA boxes
            a1 = _nhwc_empty(nt, h, w, p, dt, dev)
            _lib.call("ehgr_row_apply", ctypes.byref(op_affine(raw1, vec1[0], vec1[1], 2)), 0, a1.data_ptr(), m, p, code, sp,
                      algo_bytes=2 * m * p * es)
            # conv2 (3x3, stride s) as an implicit GEMM over the im2col operand
            pk = _pack_conv3(w2, dt, dev)
            op2 = op_conv3(a1, None, None, 0, h, w, p, False)
            if b.stride == 1:
                h2, w2_ = h, w
                raw2 = _nhwc_empty(nt, h, w, p, dt, dev)
                vec2, tr2 = bn_forward(lambda st: gemm(op2, pk[0][0], pk[0][1], raw2, st, m, 9 * p, p, "[conv3x3]", m * p), b.bn2, g2, b2, m)
            else:
                h2, w2_ = _half(h), _half(w)
                full = _nhwc_empty(nt, h, w, p, dt, dev)
                gemm(op2, pk[0][0], pk[0][1], full, 0, m, 9 * p, p, "[conv3x3]", m * p)
                raw2 = _nhwc_empty(nt, h2, w2_, p, dt, dev)
                vec2, tr2 = bn_forward(lambda st: _lib.call(
                    "ehgr_subsample2_fwd", full.data_ptr(), raw2.data_ptr(), st, nt, h, w, p, code, sp, algo_bytes=2 * raw2.numel() * es),
                    b.bn2, g2, b2, nt * h2 * w2_)
                del full
            m2 = nt * h2 * w2_
            # conv3 (1x1): bn2 + ReLU applied while the GEMM loads its operand
            raw3 = _nhwc_empty(nt, h2, w2_, cout, dt, dev)
            vec3, tr3 = bn_forward(lambda st: gemm(op_affine(raw2, vec2[0], vec2[1], 2), w3, _bf16_mirror(w3, dt, mirrors), raw3, st, m2,
                                                   p, cout, algo_in=m2 * p), b.bn3, g3, b3, m2)
            # shortcut
            xs = rawd = vecd = trd = None
            if b.down is not None:
                if b.stride == 1:
                    xs = x_blk
                else:
                    xs = _nhwc_empty(nt, h2, w2_, cin, dt, dev)
                    _lib.call("ehgr_subsample2_fwd", x_blk.data_ptr(), xs.data_ptr(), 0, nt, h, w, cin, code, sp,
                              algo_bytes=2 * xs.numel() * es)
                rawd = _nhwc_empty(nt, h2, w2_, cout, dt, dev)
                vecd, trd = bn_forward(lambda st: gemm(op_plain(xs), wd, _bf16_mirror(wd, dt, mirrors), rawd, st, m2, cin, cout,
                                                       algo_in=m2 * cin), b.down[1], gd, bd, m2)
                idn = _nhwc_empty(nt, h2, w2_, cout, dt, dev)
                _lib.call("ehgr_row_apply", ctypes.byref(op_affine(rawd, vecd[0], vecd[1], 0)), 0, idn.data_ptr(), m2, cout, code, sp,
                          algo_bytes=2 * m2 * cout * es)
            else:
                idn = x_blk
            out = _nhwc_empty(nt, h2, w2_, cout, dt, dev)
            _lib.call("ehgr_bn_add_relu", raw3.data_ptr(), vec3[0].data_ptr(), vec3[1].data_ptr(), idn.data_ptr(), out.data_ptr(), m2,
                      cout, code, sp, algo_bytes=3 * m2 * cout * es)
            del idn
            # what backward keeps is an alias of the output, never the Function output object itself
            # (output -> grad_fn -> ctx -> saved output would be a reference cycle that only the cycle GC frees)
            cur = out.detach()
            saved.append((x_blk, (h, w), raw1, vec1, tr1, a1, pk[1], raw2, vec2, tr2, raw3, vec3, tr3, xs, rawd, vecd, trd, cur))
            h, w = h2, w2_
            if b.last_of_stage and (b.stage in taps or b.stage == 4):
                outputs.append(out)
                tap_blocks.append(bi)
        if nbt:
            torch._foreach_add_(nbt, 1)
        ctx.plan, ctx.dt, ctx.params, ctx.x_in = plan, dt, params, x_in
        ctx.stem_saved, ctx.saved, ctx.tap_blocks, ctx.mirrors = stem_saved, saved, tap_blocks, mirrors
        return tuple(outputs)

    @staticmethod
    def backward(ctx, *gouts):
        plan, dt, params, x_in = ctx.plan, ctx.dt, ctx.params, ctx.x_in
        dev = x_in.device
        sp = _lib.stream_ptr(dev)
        eng = fused._STATE["engine"]
        code = _lib.F32 if dt == torch.float32 else _lib.BF16
        es = 4 if dt == torch.float32 else 2
        nt, _, H, W = x_in.shape
        # parameter gradients: straight into the gradient sink's flat buffer when one is active (train_step.GradBuckets),
        # else one zero-initialised flat buffer whose views are returned to autograd
        sink = fused._STATE["grad_sink"]
        sunk = [sink.view_for(p) if sink is not None else None for p in params]
        sizes = [p.numel() for p in params]
        gflat = torch.zeros(sum((n + 7) // 8 * 8 for n, sv in zip(sizes, sunk) if sv is None), dtype=torch.float32, device=dev)
        gviews, off = [], 0
        for p, n, sv in zip(params, sizes, sunk):
            if sv is not None:
                gviews.append(sv)
                continue
            gviews.append(gflat[off:off + n].view(p.shape))
            off += (n + 7) // 8 * 8
        bns = [plan.bn1] + [m for b in plan.blocks for m in b.modules() if isinstance(m, nn.BatchNorm2d)]
        sum_arena = _Arena(sum(2 * bn.num_features for bn in bns), torch.float64, dev, True)
        coef_arena = _Arena(sum(3 * bn.num_features for bn in bns), torch.float32, dev, False)

        def bn_backward(g, raw, vec, tr, act, gamma, ggam, gbet, count, materialise=True):
            """BatchNorm(+ReLU) backward of one layer: reduction, coefficients, d(gamma), d(beta); returns d(raw) as a tensor
            (dgrad and wgrad both read it) or as the BNBWD row operand (+ the tensors it points into)."""
            c = raw.shape[1]
            sums = sum_arena.take(2 * c)
            coef = coef_arena.take(3 * c).view(3, c)
            _lib.call("ehgr_bn_bwd_reduce_fin", g.data_ptr(), raw.data_ptr(), vec[0].data_ptr(), vec[1].data_ptr(), int(act),
                      sums.data_ptr(), count, c, code, None, sp, algo_bytes=2 * count * c * es)
            _lib.call("ehgr_bn_bwd_finalize", sums.data_ptr(), count, gamma.data_ptr(), vec[2].data_ptr(), vec[3].data_ptr(), int(tr),
                      coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(), ggam.data_ptr(), gbet.data_ptr(), c, sp)
            op = op_bnbwd(g, raw, coef[0], coef[1], coef[2], vec[0], vec[1], act)
            if not materialise:
                return op
            draw = torch.empty_like(raw)
            _lib.call("ehgr_row_apply", ctypes.byref(op), 0, draw.data_ptr(), count, c, code, sp, algo_bytes=3 * count * c * es)
            return draw

        def wgrad(dy, a_op, gw, m, k, n, tag=""):
            _lib.call("ehgr_pw_wgrad", ctypes.byref(op_plain(dy)), ctypes.byref(a_op), gw.data_ptr(), m, k, n, code, eng, sp, tag=tag,
                      algo_bytes=m * (k + n) * es + k * n * 4, algo_flops=2 * m * k * n)

        def dgrad(dy_op, w32, w16, w_is_kn, out, addend, m, k, n, tag=""):
            _lib.call("ehgr_pw_gemm_w16", ctypes.byref(dy_op), w32.data_ptr(), _lib.ptr(w16), w_is_kn, out.data_ptr(), _lib.ptr(addend),
                      0, m, k, n, code, eng, sp, tag=tag, algo_bytes=m * (k + n) * es + k * n * es, algo_flops=2 * m * k * n)

        gout_of = {bi: g for bi, g in zip(ctx.tap_blocks, gouts) if g is not None}
        # parameter index of every block (forward order)
        p_begin, p_i = [], 3
        for b in plan.blocks:
            p_begin.append(p_i)
            p_i += 9 + (3 if b.down is not None else 0)

        g = None
        for bi in range(len(plan.blocks) - 1, -1, -1):
            b = plan.blocks[bi]
            (x_blk, (h, w), raw1, vec1, tr1, a1, pk_d, raw2, vec2, tr2, raw3, vec3, tr3, xs, rawd, vecd, trd, out) = ctx.saved[bi]
            pb = p_begin[bi]
            cin, p, cout = b.conv1.in_channels, b.conv1.out_channels, b.conv3.out_channels
            h2, w2_ = out.shape[2], out.shape[3]
            m, m2 = nt * h * w, nt * h2 * w2_
            gt = gout_of.get(bi)
            if gt is not None:
                gt = fused._as_nhwc(gt, dt)
                if g is None:
                    g = gt
                else:
                    tmp = torch.empty_like(g)
                    _lib.call("ehgr_row_apply", ctypes.byref(op_plain(g)), gt.data_ptr(), tmp.data_ptr(), m2, cout, code, sp,
                              algo_bytes=3 * g.numel() * es)
                    g = tmp
            if g is None:
                raise RuntimeError("ResNet backward reached a block without an incoming gradient")
            # relu(bn3 + identity): mask by the stored output
            gz = torch.empty_like(out)
            _lib.call("ehgr_relu_bwd", g.data_ptr(), out.data_ptr(), gz.data_ptr(), gz.numel(), code, sp, algo_bytes=3 * gz.numel() * es)
            # conv3 (1x1) <- bn3 (no activation of its own)
            w3 = params[pb + 6]
            draw3 = bn_backward(gz, raw3, vec3, tr3, 0, params[pb + 7], gviews[pb + 7], gviews[pb + 8], m2)
            wgrad(draw3, op_affine(raw2, vec2[0], vec2[1], 2), gviews[pb + 6], m2, p, cout)
            g2 = _nhwc_empty(nt, h2, w2_, p, dt, dev)
            dgrad(op_plain(draw3), w3, _bf16_mirror(w3, dt, ctx.mirrors), 1, g2, None, m2, cout, p)
            del draw3
            # conv2 (3x3, stride s) <- bn2 + ReLU
            draw2 = bn_backward(g2, raw2, vec2, tr2, 2, params[pb + 4], gviews[pb + 4], gviews[pb + 5], m2)
            del g2
            if b.stride == 1:
                dfull = draw2
            else:                       # adjoint of the even-pixel sampling: zero insertion
                dfull = _nhwc_empty(nt, h, w, p, dt, dev)
                _lib.call("ehgr_subsample2_bwd", draw2.data_ptr(), dfull.data_ptr(), nt, h, w, p, code, sp,
                          algo_bytes=(draw2.numel() + dfull.numel()) * es)
            dwp = torch.zeros(p * 9 * p, dtype=torch.float32, device=dev)
            wgrad(dfull, op_conv3(a1, None, None, 0, h, w, p, False), dwp, m, 9 * p, p, "[conv3x3]")
            _lib.call("ehgr_conv3_unpack_grad", dwp.data_ptr(), gviews[pb + 3].data_ptr(), p, p, sp)
            g1 = _nhwc_empty(nt, h, w, p, dt, dev)
            dgrad(op_conv3(dfull, None, None, 0, h, w, p, False), pk_d[0], pk_d[1], 0, g1, None, m, 9 * p, p, "[conv3x3]")
            del dfull, draw2
            # shortcut gradient
            if b.down is not None:
                wdn = params[pb + 9]
                drawd = bn_backward(gz, rawd, vecd, trd, 0, params[pb + 10], gviews[pb + 10], gviews[pb + 11], m2)
                wgrad(drawd, op_plain(xs), gviews[pb + 9], m2, cin, cout)
                gxs = _nhwc_empty(nt, h2, w2_, cin, dt, dev)
                dgrad(op_plain(drawd), wdn, _bf16_mirror(wdn, dt, ctx.mirrors), 1, gxs, None, m2, cout, cin)
                del drawd
                if b.stride == 1:
                    skip = gxs
                else:
                    skip = _nhwc_empty(nt, h, w, cin, dt, dev)
                    _lib.call("ehgr_subsample2_bwd", gxs.data_ptr(), skip.data_ptr(), nt, h, w, cin, code, sp,
                              algo_bytes=(gxs.numel() + skip.numel()) * es)
            else:
                skip = gz
            # conv1 (1x1, temporal shift on its input) <- bn1 + ReLU
            w1 = params[pb]
            draw1 = bn_backward(g1, raw1, vec1, tr1, 2, params[pb + 1], gviews[pb + 1], gviews[pb + 2], m)
            del g1
            a_op = op_shift(x_blk, b.shift[0], b.shift[1], h * w, 1) if b.shift is not None else op_plain(x_blk)
            wgrad(draw1, a_op, gviews[pb], m, cin, p)
            gx = _nhwc_empty(nt, h, w, cin, dt, dev)
            if b.shift is None:         # "+ shortcut gradient" folded into the dgrad epilogue
                dgrad(op_plain(draw1), w1, _bf16_mirror(w1, dt, ctx.mirrors), 1, gx, skip, m, p, cin)
                g = gx
            else:                       # undo the shift (its adjoint) and add the shortcut gradient in one pass
                dgrad(op_plain(draw1), w1, _bf16_mirror(w1, dt, ctx.mirrors), 1, gx, None, m, p, cin)
                g = torch.empty_like(gx)
                _lib.call("ehgr_row_apply", ctypes.byref(op_shift(gx, b.shift[0], b.shift[1], h * w, -1)), skip.data_ptr(), g.data_ptr(),
                          m, cin, code, sp, algo_bytes=3 * gx.numel() * es)
            del draw1, skip, gz
            if sink is not None:
                n_p = 9 + (3 if b.down is not None else 0)
                sink.mark_done([q for q, sv in zip(params[pb:pb + n_p], sunk[pb:pb + n_p]) if sv is not None])
        # ---- stem: max-pool adjoint -> bn1 + ReLU backward (BNBWD operand) -> 7x7 weight gradient
        raw0, vec0, tr0, pool_idx, (ho, wo), patches = ctx.stem_saved
        gt = gout_of.get(-1)              # gradient of the tapped max-pool output
        if gt is not None:
            gt = fused._as_nhwc(gt, dt)
            tmp = torch.empty_like(g)
            _lib.call("ehgr_row_apply", ctypes.byref(op_plain(g)), gt.data_ptr(), tmp.data_ptr(), g.numel() // 64, 64, code, sp,
                      algo_bytes=3 * g.numel() * es)
            g = tmp
        g_act = torch.empty_like(raw0)
        _lib.call("ehgr_maxpool3_bwd", g.data_ptr(), pool_idx.data_ptr(), g_act.data_ptr(), nt, ho, wo, 64, code, sp,
                  algo_bytes=g.numel() * (es + 1) + g_act.numel() * es)
        if patches is not None:
            m0 = nt * ho * wo
            draw0 = bn_backward(g_act, raw0, vec0, tr0, 2, params[1], gviews[1], gviews[2], m0)
            del g_act
            dwp = torch.zeros(64 * STEM_KP, dtype=torch.float32, device=dev)
            wgrad(draw0, op_plain(patches), dwp, m0, STEM_KP, 64, "[stem7x7]")
            _lib.call("ehgr_stem7_unpack_grad", dwp.data_ptr(), gviews[0].data_ptr(), 64, STEM_KP, sp)
        else:
            dy_op = bn_backward(g_act, raw0, vec0, tr0, 2, params[1], gviews[1], gviews[2], nt * ho * wo, materialise=False)
            _lib.call("ehgr_stem7_wgrad", ctypes.byref(dy_op), x_in.data_ptr(), gviews[0].data_ptr(), nt, H, W, 64,
                      _lib.dtype_code(x_in), code, sp, algo_bytes=2 * raw0.numel() * es + x_in.numel() * x_in.element_size(),
                      algo_flops=2 * 147 * raw0.numel())
        if sink is not None:
            sink.mark_done([q for q, sv in zip(params[0:3], sunk[0:3]) if sv is not None])
        pg = [gv if (q.requires_grad and sv is None) else None for gv, q, sv in zip(gviews, params, sunk)]
        return (None, None, None, None, *pg)


def resnet_features(model, x, taps: Sequence[int] = ()):
    """conv1 .. layer4 of a torchvision Bottleneck ResNet -> [NT, 2048, H/32, W/32] (NHWC strides, compute dtype).  With
    ``taps`` (stage numbers among 1, 2, 3; 0 = the max-pool output) returns a tuple: those outputs (ascending) followed by
    the layer4 output."""
    _lib.require_cuda(x)
    plan = plan_of(model)
    taps = tuple(sorted(set(int(t) for t in taps)))
    if any(t not in (0, 1, 2, 3) for t in taps):
        raise ValueError("taps are stage numbers among 1, 2, 3 (0 = the max-pool output)")
    outs = _ResNetFunction.apply(plan, fused._STATE["dtype"], taps, x, *_plan_params(plan))
    return outs[0] if not taps else outs
