"""Loss heads of the two training stages — one fused forward+backward kernel each (csrc/loss.cu).

* ``mtmm_loss``  — train_mtmm.py:223-231: CE(logits, labels) + 0.01 * MSE(depth_pred,
  bilinear56(depth_gt)).  The 224->56 bilinear resize with align_corners=False is exactly the mean of
  the 2x2 pixels at rows/cols 4i+1, 4i+2, which is what the kernel reads directly.
* ``sd_loss``    — train_sd.py:178-193,227-265: 4 CE + 3 temperature-KL (x T^2) + 3 masked-L2 feature
  terms, weights (1-alpha), alpha, beta.

Both return device scalars (no .item() synchronisation: the reference syncs 4 / 17 times per step).
The gradients w.r.t. logits / depth prediction / features are produced by the same launch and
handed to autograd in ``backward`` (scaled by the incoming gradient of the loss).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class _MTMMLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, pred, labels, depth_gt, depth_weight):
        _lib.require_cuda(logits, pred, labels, depth_gt)
        n, k = logits.shape
        ph, pw = pred.shape[-2:]
        frames = pred.numel() // (ph * pw)
        if depth_gt.shape[-2] != 4 * ph or depth_gt.shape[-1] != 4 * pw or depth_gt.numel() != frames * 16 * ph * pw:
            raise RuntimeError("mtmm_loss expects depth_gt of 4x the spatial size of depth_pred and the same frame count")
        lg = logits.contiguous().float()
        pr = pred.contiguous()
        if pr.dtype not in (torch.float32, torch.bfloat16):
            pr = pr.float()
        gt = depth_gt.contiguous().float()
        lab = labels.contiguous().long()
        out = torch.zeros(3, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(lg)
        dpred = torch.empty(pr.shape, dtype=torch.float32, device=pr.device)
        _lib.call("ehgr_mtmm_loss", lg.data_ptr(), lab.data_ptr(), pr.data_ptr(), gt.data_ptr(), float(depth_weight),
                  out.data_ptr(), dlogits.data_ptr(), dpred.data_ptr(), n, k, frames, ph, pw, _lib.dtype_code(pr),
                  _lib.stream_ptr(logits.device), algo_bytes=gt.numel() * 2 + pr.numel() * (pr.element_size() + 4))
        ctx.save_for_backward(dlogits, dpred)
        ctx.dtypes = (logits.dtype, pred.dtype)
        return out                                  # [total, CE, MSE]; only out[0] carries gradient

    @staticmethod
    def backward(ctx, g_out):
        dlogits, dpred = ctx.saved_tensors
        g_total = g_out[0]
        return (dlogits * g_total).to(ctx.dtypes[0]), (dpred * g_total).to(ctx.dtypes[1]), None, None, None


def mtmm_loss(logits, labels, depth_pred, depth_gt, depth_weight: float = 0.01):
    """Returns (loss, depth_mse) as device scalars.  depth_gt: [..., 4*ph, 4*pw] fp32 in [0,1]."""
    out = _MTMMLossFunction.apply(logits, depth_pred, labels, depth_gt, depth_weight)
    return out[0], out[2].detach()


class _SDLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels, alpha, beta, temperature, *tensors):
        logits, feats = tensors[:4], tensors[4:]
        _lib.require_cuda(*tensors)
        n, k = logits[0].shape
        lg = [t.contiguous().float() for t in logits]
        ft = [t.contiguous().float().reshape(t.shape[0], -1) for t in feats]
        rows, f = ft[0].shape
        if any(t.shape != (rows, f) for t in ft) or any(t.shape != (n, k) for t in lg):
            raise RuntimeError("sd_loss: the four logits / four feature tensors must have equal shapes")
        lab = labels.contiguous().long()
        terms = torch.zeros(11, dtype=torch.float32, device=lab.device)
        dl = [torch.empty_like(t) for t in lg]
        df = [torch.empty_like(t) for t in ft[1:]]
        P4, P3 = ctypes.c_void_p * 4, ctypes.c_void_p * 3
        _lib.call("ehgr_sd_loss", P4(*[t.data_ptr() for t in lg]), P4(*[t.data_ptr() for t in ft]), lab.data_ptr(),
                  float(alpha), float(beta), float(temperature), terms.data_ptr(), P4(*[t.data_ptr() for t in dl]),
                  P3(*[t.data_ptr() for t in df]), n, k, rows, f, _lib.stream_ptr(lab.device),
                  algo_bytes=7 * rows * f * 4)
        ctx.save_for_backward(*dl, *df)
        ctx.meta = ([t.dtype for t in logits], [t.dtype for t in feats], [t.shape for t in feats])
        return terms                                # [total, 4 CE, 3 KD, 3 feature]; only terms[0] carries gradient

    @staticmethod
    def backward(ctx, g_terms):
        saved = ctx.saved_tensors
        g_total = g_terms[0]
        dl, df = saved[:4], saved[4:]
        ldt, fdt, fshape = ctx.meta
        gl = [(d * g_total).to(t) for d, t in zip(dl, ldt)]
        gf = [None] + [(d * g_total).to(t).reshape(s) for d, t, s in zip(df, fdt[1:], fshape[1:])]
        return (None, None, None, None, *gl, *gf)


def sd_loss(outputs, feats, labels, alpha: float = 0.1, beta: float = 1e-6, temperature: float = 3.0):
    """outputs = (final, mid1, mid2, mid3) logits [N,cls]; feats = (final, mid1, mid2, mid3) pooled features
    (any trailing shape).  Returns (total, terms[10]): 4 CE, 3 KD (already x T^2), 3 feature sums."""
    if len(outputs) != 4 or len(feats) != 4:
        raise ValueError("sd_loss expects four logits and four feature tensors (final first)")
    terms = _SDLossFunction.apply(labels, alpha, beta, temperature, *outputs, *feats)
    return terms[0], terms[1:].detach()


class _DepthMSEFunction(torch.autograd.Function):
    """weight * MSE(pred, bilinear56(depth_gt)) alone (ehgr_mtmm_loss with n = 0)."""

    @staticmethod
    def forward(ctx, pred, depth_gt, weight):
        _lib.require_cuda(pred, depth_gt)
        ph, pw = pred.shape[-2:]
        frames = pred.numel() // (ph * pw)
        if depth_gt.shape[-2] != 4 * ph or depth_gt.shape[-1] != 4 * pw or depth_gt.numel() != frames * 16 * ph * pw:
            raise RuntimeError("depth loss expects depth_gt of 4x the spatial size of depth_pred and the same frame count")
        pr = pred.contiguous()
        if pr.dtype not in (torch.float32, torch.bfloat16):
            pr = pr.float()
        gt = depth_gt.contiguous().float()
        out = torch.zeros(3, dtype=torch.float32, device=pred.device)
        dpred = torch.empty(pr.shape, dtype=torch.float32, device=pr.device)
        _lib.call("ehgr_mtmm_loss", 0, 0, pr.data_ptr(), gt.data_ptr(), float(weight), out.data_ptr(), 0, dpred.data_ptr(),
                  0, 0, frames, ph, pw, _lib.dtype_code(pr), _lib.stream_ptr(pred.device),
                  algo_bytes=gt.numel() * 2 + pr.numel() * (pr.element_size() + 4))
        ctx.save_for_backward(dpred)
        ctx.dtype = pred.dtype
        return out                                  # [weight * MSE, 0, MSE]

    @staticmethod
    def backward(ctx, g_out):
        (dpred,) = ctx.saved_tensors
        return (dpred * g_out[0]).to(ctx.dtype), None, None


def mtmm_sd_loss(outputs, feats, global_depth_out, depth_gt, labels, alpha: float = 0.1, beta: float = 1e-6,
                 temperature: float = 3.0, depth_weight: float = 0.01):
    """The combined stage of train_mtmm_sd.py:240-293: ``loss = CE(output) + 0.01 * MSE(g_depth_out, bilinear56(depth))``
    and ``total = (1-alpha) * (loss + 3 CE) + alpha * 3 KD + beta * 3 feature`` — the self-distillation kernel's total plus
    ``(1-alpha) * 0.01 * MSE`` from the depth kernel (two launches; both return their gradients with the forward).
    Returns (total, terms[10] of sd_loss, depth_mse)."""
    total_sd, terms = sd_loss(outputs, feats, labels, alpha, beta, temperature)
    d = _DepthMSEFunction.apply(global_depth_out, depth_gt, (1.0 - alpha) * depth_weight)
    return total_sd + d[0], terms, d[2].detach()
