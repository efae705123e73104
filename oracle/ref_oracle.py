"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A plain PyTorch-fp32 / numpy restatement of the reference's algorithm for the training hot path
(TSM/ACTION-MobileNetV2 forward, TSN wrapper, MTMM and SD loss heads).  It is functional code over a
``state_dict`` (name -> tensor, the reference's parameter names) so that the same weights can be
fed to the reference modules, to this oracle and to the CUDA implementation.  Gradients come from
autograd over these functions.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker or the reported CPU baseline — the product
package never imports it.

Parity pinning: ``tests/golden/make_golden.py`` runs the live reference modules (``/root/reference``,
present in the build container only), asserts there that this file agrees with them, and stores the
reference outputs as fixtures that travel to the GPU box; ``tests/test_oracle_golden.py`` checks the
oracle against those fixtures everywhere.  Pinned that way: shift, Action, the whole TSN-MobileNetV2
(none / TSM / ACTION x train / eval BN), both loss heads, the MTMM ``global_decoder``, ``SepConv``, the
MTMM+SD decoders and combined loss, ``EMAWrapper`` and ``TemporalPool``.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# archs/mobilenet_v2.py:75-84 — (t, c, n, s)
MBV2_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                (6, 320, 1, 1))


def mbv2_block_table():
    """[(index in features, inp, oup, stride, expand_ratio)] for the 17 InvertedResidual blocks."""
    rows, inp, idx = [], 32, 1
    for t, c, n, s in MBV2_SETTING:
        for i in range(n):
            rows.append((idx, inp, c, s if i == 0 else 1, t))
            inp, idx = c, idx + 1
    return rows


# --------------------------------------------------------------------------------------------
# K1 temporal shift — models/temporal_shift.py:27-46 (dup. models/action.py:135-154)
# --------------------------------------------------------------------------------------------
def temporal_shift_np(x: np.ndarray, n_segment: int, fold_div: int) -> np.ndarray:
    """numpy, any dtype (bit-exact copy semantics).  x: [nt, c, h, w]."""
    nt, c, h, w = x.shape
    n_batch = nt // n_segment
    v = x.reshape(n_batch, n_segment, c, h, w)  # raises like the reference's .view on a bad nt
    fold = c // fold_div
    out = np.zeros_like(v)
    out[:, :-1, :fold] = v[:, 1:, :fold]
    out[:, 1:, fold:2 * fold] = v[:, :-1, fold:2 * fold]
    out[:, :, 2 * fold:] = v[:, :, 2 * fold:]
    return out.reshape(nt, c, h, w)


def temporal_shift_bwd_np(g: np.ndarray, n_segment: int, fold_div: int) -> np.ndarray:
    """Adjoint of temporal_shift_np (what autograd produces; equals InplaceShift.backward,
    models/temporal_shift.py:64-76)."""
    nt, c, h, w = g.shape
    n_batch = nt // n_segment
    v = g.reshape(n_batch, n_segment, c, h, w)
    fold = c // fold_div
    out = np.zeros_like(v)
    out[:, 1:, :fold] = v[:, :-1, :fold]
    out[:, :-1, fold:2 * fold] = v[:, 1:, fold:2 * fold]
    out[:, :, 2 * fold:] = v[:, :, 2 * fold:]
    return out.reshape(nt, c, h, w)


def temporal_shift(x: torch.Tensor, n_segment: int, fold_div: int) -> torch.Tensor:
    """torch, differentiable."""
    nt, c, h, w = x.shape
    v = x.view(nt // n_segment, n_segment, c, h, w)
    fold = c // fold_div
    z = torch.zeros_like(v[:, :1])
    left = torch.cat([v[:, 1:, :fold], z[:, :, :fold]], 1)
    right = torch.cat([z[:, :, fold:2 * fold], v[:, :-1, fold:2 * fold]], 1)
    return torch.cat([left, right, v[:, :, 2 * fold:]], 2).reshape(nt, c, h, w)


# --------------------------------------------------------------------------------------------
# BatchNorm helper (nn.BatchNorm2d semantics: biased var to normalise, unbiased for running stat)
# --------------------------------------------------------------------------------------------
def _bn(x, sd: SD, prefix: str, training: bool, momentum: float = 0.1, eps: float = 1e-5):
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, momentum, eps)


# --------------------------------------------------------------------------------------------
# ACTION — models/action.py:61-116, op for op (not the algebraic collapse)
# --------------------------------------------------------------------------------------------
def action_pre_net(x: torch.Tensor, sd: SD, prefix: str, n_segment: int, bn_training: bool) -> torch.Tensor:
    """Everything in Action.forward up to (not including) ``self.net``: returns x_p1+x_p2+x_p3."""
    nt, c, h, w = x.shape
    n = nt // n_segment
    T = n_segment
    # models/action.py:65-73 — depthwise temporal conv on (n*h*w, c, T)
    xs = x.view(n, T, c, h, w).permute(0, 3, 4, 2, 1).reshape(n * h * w, c, T)
    xs = F.conv1d(xs, sd[prefix + ".action_shift.weight"], padding=1, groups=c)
    xs = xs.view(n, h, w, c, T).permute(0, 4, 3, 1, 2).reshape(nt, c, h, w)
    # :76-83 — spatio-temporal excitation
    p1 = xs.view(n, T, c, h, w).transpose(2, 1).mean(1, keepdim=True)           # [n,1,T,h,w]
    p1 = F.conv3d(p1, sd[prefix + ".action_p1_conv1.weight"], padding=1)
    p1 = torch.sigmoid(p1.transpose(2, 1).reshape(nt, 1, h, w))
    x_p1 = xs * p1 + xs
    # :86-96 — channel excitation
    p2 = F.adaptive_avg_pool2d(xs, 1)
    p2 = F.conv2d(p2, sd[prefix + ".action_p2_squeeze.weight"])
    cr = p2.shape[1]
    p2 = p2.view(n, T, cr).transpose(2, 1)
    p2 = F.relu(F.conv1d(p2, sd[prefix + ".action_p2_conv1.weight"], padding=1))
    p2 = p2.transpose(2, 1).reshape(nt, cr, 1, 1)
    p2 = torch.sigmoid(F.conv2d(p2, sd[prefix + ".action_p2_expand.weight"]))
    x_p2 = xs * p2 + xs
    # :99-113 — motion excitation
    x3 = F.conv2d(xs, sd[prefix + ".action_p3_squeeze.weight"])
    x3 = _bn(x3, sd, prefix + ".action_p3_bn1", bn_training)
    x3_t = x3.view(n, T, cr, h, w)[:, :T - 1]
    x3_t1 = F.conv2d(x3, sd[prefix + ".action_p3_conv1.weight"], padding=1, groups=cr)
    x3_t1 = x3_t1.view(n, T, cr, h, w)[:, 1:]
    p3 = F.pad(x3_t1 - x3_t, (0, 0, 0, 0, 0, 0, 0, 1))
    p3 = F.adaptive_avg_pool2d(p3.reshape(nt, cr, h, w), 1)
    p3 = torch.sigmoid(F.conv2d(p3, sd[prefix + ".action_p3_expand.weight"]))
    x_p3 = xs * p3 + xs
    return x_p1 + x_p2 + x_p3


def action_forward(x, sd: SD, prefix: str, n_segment: int, bn_training: bool):
    """Action wrapping a bias-free 1x1 conv ``net`` (the MobileNetV2 case, models/models.py:183-185)."""
    return F.conv2d(action_pre_net(x, sd, prefix, n_segment, bn_training), sd[prefix + ".net.weight"])


# --------------------------------------------------------------------------------------------
# MobileNetV2 — archs/mobilenet_v2.py:28-66 (block), :110-114 (forward)
# --------------------------------------------------------------------------------------------
def inverted_residual(x, sd: SD, prefix: str, inp: int, oup: int, stride: int, expand: int,
                      temporal: str, n_segment: int, shift_div: int, bn_training: bool):
    """``prefix`` = 'base_model.features.{i}'.  temporal in {'none','tsm','action'} applies to conv[0]
    of residual blocks only (models/models.py:183)."""
    hidden = inp * expand
    res = stride == 1 and inp == oup
    y = x
    k = 0
    if expand != 1:
        p0 = f"{prefix}.conv.0"
        if res and temporal == "tsm":
            y = F.conv2d(temporal_shift(y, n_segment, shift_div), sd[p0 + ".net.weight"])
        elif res and temporal == "action":
            y = action_forward(y, sd, p0, n_segment, bn_training)
        else:
            y = F.conv2d(y, sd[p0 + ".weight"])
        y = F.relu6(_bn(y, sd, f"{prefix}.conv.1", bn_training))
        k = 3
    y = F.conv2d(y, sd[f"{prefix}.conv.{k}.weight"], stride=stride, padding=1, groups=hidden)
    y = F.relu6(_bn(y, sd, f"{prefix}.conv.{k + 1}", bn_training))
    y = F.conv2d(y, sd[f"{prefix}.conv.{k + 3}.weight"])
    y = _bn(y, sd, f"{prefix}.conv.{k + 4}", bn_training)
    return x + y if res else y


def mobilenet_v2_features(x, sd: SD, temporal: str = "none", n_segment: int = 8, shift_div: int = 8,
                          bn_training: bool = True, prefix: str = "base_model.features", taps=None):
    """features[0..18]; ``taps`` (optional dict) receives {index: activation} for the listed indices."""
    y = F.conv2d(x, sd[f"{prefix}.0.0.weight"], stride=2, padding=1)
    y = F.relu6(_bn(y, sd, f"{prefix}.0.1", bn_training))
    for idx, inp, oup, stride, t in mbv2_block_table():
        y = inverted_residual(y, sd, f"{prefix}.{idx}", inp, oup, stride, t, temporal, n_segment, shift_div,
                              bn_training)
        if taps is not None and idx in taps:
            taps[idx] = y
    y = F.conv2d(y, sd[f"{prefix}.18.0.weight"])
    y = F.relu6(_bn(y, sd, f"{prefix}.18.1", bn_training))
    if taps is not None and 18 in taps:
        taps[18] = y
    return y


def tsn_forward(x5: torch.Tensor, sd: SD, num_segments: int, temporal: str = "none", shift_div: int = 8,
                bn_training: bool = True, dropout_mask: Optional[torch.Tensor] = None, taps=None):
    """models/models.py:323-356 with base_model='mobilenetv2': fold T into the batch, backbone,
    x.mean(3).mean(2) (archs/mobilenet_v2.py:112), Dropout (as an explicit mask*scale tensor, or
    identity), new_fc, mean over segments.  Returns logits [N, num_class]."""
    x = x5.view((-1, 3) + tuple(x5.shape[-2:]))
    f = mobilenet_v2_features(x, sd, temporal, num_segments, shift_div, bn_training, taps=taps)
    pooled = f.mean(3).mean(2)
    if dropout_mask is not None:
        pooled = pooled * dropout_mask
    z = F.linear(pooled, sd["new_fc.weight"], sd["new_fc.bias"])
    return z.view((-1, num_segments) + tuple(z.shape[1:])).mean(dim=1, keepdim=True).squeeze(1)


# --------------------------------------------------------------------------------------------
# Loss heads
# --------------------------------------------------------------------------------------------
def mtmm_loss(logits, labels, depth_pred, depth_gt5):
    """train_mtmm.py:223-231.  depth_gt5: [N,T,1,224,224]; depth_pred: [NT,1,56,56]."""
    gt = depth_gt5.view(-1, 1, depth_gt5.size(-2), depth_gt5.size(-1))
    gt = F.interpolate(gt, size=(56, 56), mode="bilinear")
    g_depth_loss = F.mse_loss(depth_pred, gt)
    return F.cross_entropy(logits, labels) + 0.01 * g_depth_loss, g_depth_loss


def kd_loss(output, target_soft, temperature):
    """train_sd.py:178-188."""
    ls = torch.log_softmax(output / temperature, dim=1)
    return -torch.mean(torch.sum(ls * target_soft, dim=1))


def feature_loss(fea, target_fea):
    """train_sd.py:191-193."""
    loss = (fea - target_fea) ** 2 * ((fea > 0) | (target_fea > 0)).float()
    return torch.abs(loss).sum()


def sd_loss(outputs, feats, labels, alpha=0.1, beta=1e-6, temperature=3.0):
    """train_sd.py:227-265.  outputs = (final, mid1, mid2, mid3) logits; feats = (final, mid1, mid2,
    mid3) pooled features.  Returns (total, dict of the ten terms)."""
    out, m1, m2, m3 = outputs
    f4, f1, f2, f3 = feats
    ce = [F.cross_entropy(o, labels) for o in (out, m1, m2, m3)]
    temp4 = torch.softmax(out / temperature, dim=1).detach()
    kd = [kd_loss(m, temp4, temperature) * temperature ** 2 for m in (m1, m2, m3)]
    fl = [feature_loss(f, f4.detach()) for f in (f1, f2, f3)]
    total = (1 - alpha) * sum(ce) + alpha * sum(kd) + beta * sum(fl)
    return total, {"ce": ce, "kd": kd, "feat": fl}


# --------------------------------------------------------------------------------------------
# state-dict utilities
# --------------------------------------------------------------------------------------------
def clone_state(sd: SD, requires_grad: bool = True, dtype=torch.float32) -> SD:
    """Detached fp32 copy; floating-point *parameters* get requires_grad (buffers do not)."""
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if t.is_floating_point():
            t = t.to(dtype)
            if requires_grad and not (k.endswith("running_mean") or k.endswith("running_var")):
                t.requires_grad_(True)
        out[k] = t
    return out


def randomize_state(sd: SD, seed: int = 0, action_sigma: float = 0.3) -> None:
    """In-place: give BN layers non-trivial affine/running stats and ACTION layers non-trivial
    weights so that parity runs do not sit on the initialisation's symmetric point."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if not v.is_floating_point():
            continue
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        elif ".action_" in k and k.endswith("weight") and "bn" not in k:
            v.add_(torch.randn(v.shape, generator=g) * action_sigma)
        elif k.endswith("weight") and v.dim() == 1:   # BN gamma
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        elif k.endswith("bias") and v.dim() == 1 and "fc" not in k:  # BN beta
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)


# --------------------------------------------------------------------------------------------
# deterministic, machine-independent weights and inputs (numpy RandomState, not torch RNG)
# --------------------------------------------------------------------------------------------
def _bn_entries(sd, prefix, c, rs):
    sd[prefix + ".weight"] = torch.from_numpy(rs.uniform(0.75, 1.25, c).astype(np.float32))
    sd[prefix + ".bias"] = torch.from_numpy((rs.standard_normal(c) * 0.1).astype(np.float32))
    sd[prefix + ".running_mean"] = torch.from_numpy((rs.standard_normal(c) * 0.1).astype(np.float32))
    sd[prefix + ".running_var"] = torch.from_numpy(rs.uniform(0.75, 1.25, c).astype(np.float32))
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def _conv_entry(sd, name, shape, rs, std=None):
    if std is None:  # archs/mobilenet_v2.py:118-120
        std = math.sqrt(2.0 / (shape[2] * shape[3] * shape[0]))
    sd[name] = torch.from_numpy((rs.standard_normal(shape) * std).astype(np.float32))


def action_state(sd: SD, prefix: str, c: int, shift_div: int, rs, sigma: float = 0.3) -> None:
    """Entries of one Action module (models/action.py:25-58): TSM-pattern shift weights plus noise,
    default-scale random weights elsewhere."""
    cr, fold = c // 16, c // shift_div
    w = np.zeros((c, 1, 3), np.float32)
    w[:fold, 0, 2] = 1
    w[fold:2 * fold, 0, 0] = 1
    w[2 * fold:, 0, 1] = 1
    sd[prefix + ".action_shift.weight"] = torch.from_numpy(w + (rs.standard_normal(w.shape) * sigma).astype(np.float32))
    sd[prefix + ".action_p1_conv1.weight"] = torch.from_numpy((rs.standard_normal((1, 1, 3, 3, 3)) * 0.2).astype(np.float32))
    sd[prefix + ".action_p2_squeeze.weight"] = torch.from_numpy((rs.standard_normal((cr, c, 1, 1)) * 0.3).astype(np.float32))
    sd[prefix + ".action_p2_conv1.weight"] = torch.from_numpy((rs.standard_normal((cr, cr, 3)) * 0.5).astype(np.float32))
    sd[prefix + ".action_p2_expand.weight"] = torch.from_numpy((rs.standard_normal((c, cr, 1, 1)) * 0.5).astype(np.float32))
    sd[prefix + ".action_p3_squeeze.weight"] = torch.from_numpy((rs.standard_normal((cr, c, 1, 1)) * 0.3).astype(np.float32))
    _bn_entries(sd, prefix + ".action_p3_bn1", cr, rs)
    sd[prefix + ".action_p3_conv1.weight"] = torch.from_numpy((rs.standard_normal((cr, 1, 3, 3)) * 0.4).astype(np.float32))
    sd[prefix + ".action_p3_expand.weight"] = torch.from_numpy((rs.standard_normal((c, cr, 1, 1)) * 0.5).astype(np.float32))


def build_tsn_state(num_class: int = 83, temporal: str = "none", shift_div: int = 8, seed: int = 0) -> SD:
    """A full state_dict for TSN(base_model='mobilenetv2', dropout>0) with the REFERENCE's key names
    (models/models.py:169-194 + archs/mobilenet_v2.py), loadable with strict=True into the reference
    and into ehgr_b200.TSN.  Values come from numpy RandomState(seed)."""
    rs = np.random.RandomState(seed)
    sd: SD = {}
    f = "base_model.features"
    _conv_entry(sd, f"{f}.0.0.weight", (32, 3, 3, 3), rs)
    _bn_entries(sd, f"{f}.0.1", 32, rs)
    for idx, inp, oup, stride, t in mbv2_block_table():
        hidden, p, k = inp * t, f"{f}.{idx}.conv", 0
        res = stride == 1 and inp == oup
        if t != 1:
            if res and temporal in ("tsm", "action"):
                _conv_entry(sd, f"{p}.0.net.weight", (hidden, inp, 1, 1), rs)
                if temporal == "action":
                    action_state(sd, f"{p}.0", inp, shift_div, rs)
            else:
                _conv_entry(sd, f"{p}.0.weight", (hidden, inp, 1, 1), rs)
            _bn_entries(sd, f"{p}.1", hidden, rs)
            k = 3
        _conv_entry(sd, f"{p}.{k}.weight", (hidden, 1, 3, 3), rs)
        _bn_entries(sd, f"{p}.{k + 1}", hidden, rs)
        _conv_entry(sd, f"{p}.{k + 3}.weight", (oup, hidden, 1, 1), rs)
        _bn_entries(sd, f"{p}.{k + 4}", oup, rs)
    _conv_entry(sd, f"{f}.18.0.weight", (1280, 320, 1, 1), rs)
    _bn_entries(sd, f"{f}.18.1", 1280, rs)
    sd["new_fc.weight"] = torch.from_numpy((rs.standard_normal((num_class, 1280)) * 0.02).astype(np.float32))
    sd["new_fc.bias"] = torch.from_numpy((rs.standard_normal(num_class) * 0.02).astype(np.float32))
    return sd


def synthetic_clip_batch(n: int, t: int = 8, size: int = 224, num_class: int = 83, seed: int = 0):
    """(rgb [n,t,3,size,size] ~N(0,1), depth [n,t,1,size,size] in [0,1], labels [n]) — SURVEY §8d."""
    rs = np.random.RandomState(1000 + seed)
    rgb = torch.from_numpy(rs.standard_normal((n, t, 3, size, size)).astype(np.float32))
    depth = torch.from_numpy(rs.uniform(0, 1, (n, t, 1, size, size)).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, num_class, (n,)).astype(np.int64))
    return rgb, depth, labels


# --------------------------------------------------------------------------------------------
# MTMM wrapper on MobileNetV2 — models/models_MTMM.py:129-155 (global_decoder), :268-292 (forward)
# The reference wires this for ResNet only (layer4, 2048 ch); for MobileNetV2 the tap is the
# features[18] output (1280 ch) — builder-defined, SURVEY §8a A10.
# --------------------------------------------------------------------------------------------
DECODER_UNITS = ((0, 1, True), (4, 5, True), (8, 9, True), (12, 13, False))  # (conv idx, bn idx, upsample)


def global_decoder(f, sd: SD, bn_training: bool, prefix: str = "global_decoder"):
    y = f
    for ci, bi, up in DECODER_UNITS:
        y = F.conv2d(y, sd[f"{prefix}.{ci}.weight"], padding=1)
        y = F.relu(_bn(y, sd, f"{prefix}.{bi}", bn_training))
        if up:
            y = F.interpolate(y, scale_factor=2, mode="nearest")
    return torch.sigmoid(F.conv2d(y, sd[f"{prefix}.15.weight"], sd[f"{prefix}.15.bias"]))


def mtmm_forward(x5, sd: SD, num_segments: int, temporal: str = "tsm", shift_div: int = 8,
                 bn_training: bool = True, dropout_mask=None):
    """-> (logits [N,cls], depth [NT,1,56,56])."""
    taps = {18: None}
    logits = tsn_forward(x5, sd, num_segments, temporal, shift_div, bn_training, dropout_mask, taps=taps)
    return logits, global_decoder(taps[18], sd, bn_training)


def decoder_state(sd: SD, rs, feat: int = 1280, prefix: str = "global_decoder") -> None:
    """Entries of the MTMM depth decoder (models/models_MTMM.py:129-155) for `feat` input channels."""
    chans = (feat, 256, 64, 32, 32)
    for (ci, bi, _), i, o in zip(DECODER_UNITS, chans[:-1], chans[1:]):
        sd[f"{prefix}.{ci}.weight"] = torch.from_numpy(
            (rs.standard_normal((o, i, 3, 3)) * math.sqrt(2.0 / (9 * i))).astype(np.float32))
        _bn_entries(sd, f"{prefix}.{bi}", o, rs)
    sd[f"{prefix}.15.weight"] = torch.from_numpy((rs.standard_normal((1, 32, 1, 1)) * 0.2).astype(np.float32))
    sd[f"{prefix}.15.bias"] = torch.from_numpy((rs.standard_normal(1) * 0.1).astype(np.float32))


def build_mtmm_state(num_class: int = 83, temporal: str = "tsm", shift_div: int = 8, seed: int = 0,
                     feat: int = 1280) -> SD:
    sd = build_tsn_state(num_class, temporal, shift_div, seed)
    decoder_state(sd, np.random.RandomState(seed + 77), feat)
    return sd


def mtmm_train_step(sd: SD, rgb, depth, labels, num_segments=8, temporal="tsm", shift_div=8, bn_training=True):
    """One reference-equivalent MTMM step on CPU: forward, loss (train_mtmm.py:223-231), backward.
    Returns (loss, logits, depth_pred); gradients are left in sd[*].grad."""
    for v in sd.values():
        if v.requires_grad:
            v.grad = None
    logits, dpred = mtmm_forward(rgb, sd, num_segments, temporal, shift_div, bn_training)
    loss, _ = mtmm_loss(logits, labels, dpred, depth)
    loss.backward()
    return loss.detach(), logits.detach(), dpred.detach()


# --------------------------------------------------------------------------------------------
# SD wrapper on MobileNetV2 — exit heads models/models_SD.py:81-101 (SepConv), :214-253 (scala1-3,
# middle_fc1-3), forward :364-431.  The reference wires ResNet stages; the MobileNetV2 taps
# (features[3], [6], [13]; 24/32/96 channels) are builder-defined.
# --------------------------------------------------------------------------------------------
def sepconv(x, sd: SD, prefix: str, bn_training: bool):
    c = x.shape[1]
    y = F.conv2d(x, sd[f"{prefix}.op.0.weight"], stride=2, padding=1, groups=c)
    y = F.conv2d(y, sd[f"{prefix}.op.1.weight"])
    y = F.relu(_bn(y, sd, f"{prefix}.op.2", bn_training))
    y = F.conv2d(y, sd[f"{prefix}.op.4.weight"], stride=1, padding=1, groups=c)
    y = F.conv2d(y, sd[f"{prefix}.op.5.weight"])
    return F.relu(_bn(y, sd, f"{prefix}.op.6", bn_training))


SD_TAPS = (3, 6, 13)
SD_HEADS = (("scala1", (24, 32, 96, 1280)), ("scala2", (32, 96, 1280)), ("scala3", (96, 1280)))


def sd_forward(x5, sd: SD, num_segments: int, temporal: str = "tsm", shift_div: int = 8, bn_training: bool = True):
    """-> (output, mid1, mid2, mid3, final_fea, fea1, fea2, fea3) as models/models_SD.py:431."""
    taps = {i: None for i in SD_TAPS + (18,)}
    x = x5.view((-1, 3) + tuple(x5.shape[-2:]))
    mobilenet_v2_features(x, sd, temporal, num_segments, shift_div, bn_training, taps=taps)
    mids, feas = [], []
    for (name, chans), tap, fc in zip(SD_HEADS, SD_TAPS, ("middle_fc1", "middle_fc2", "middle_fc3")):
        y = taps[tap]
        for j in range(len(chans) - 1):
            y = sepconv(y, sd, f"{name}.{j}", bn_training)
        fea = F.adaptive_avg_pool2d(y, 1)
        z = F.linear(torch.flatten(fea, 1), sd[fc + ".weight"], sd[fc + ".bias"])
        mids.append(z.view((-1, num_segments) + tuple(z.shape[1:])).mean(1))
        feas.append(fea)
    final_fea = F.adaptive_avg_pool2d(taps[18], 1)
    z = F.linear(torch.flatten(final_fea, 1), sd["new_fc.weight"], sd["new_fc.bias"])
    out = z.view((-1, num_segments) + tuple(z.shape[1:])).mean(1)
    return (out, *mids, final_fea, *feas)


def sepconv_state(sd: SD, prefix: str, ci: int, co: int, rs) -> None:
    """Entries of one SepConv (models/models_SD.py:81-101): `prefix`.op.{0,1,2,4,5,6}."""
    p = f"{prefix}.op"
    _conv_entry(sd, f"{p}.0.weight", (ci, 1, 3, 3), rs, std=0.3)
    _conv_entry(sd, f"{p}.1.weight", (ci, ci, 1, 1), rs, std=math.sqrt(2.0 / ci))
    _bn_entries(sd, f"{p}.2", ci, rs)
    _conv_entry(sd, f"{p}.4.weight", (ci, 1, 3, 3), rs, std=0.3)
    _conv_entry(sd, f"{p}.5.weight", (co, ci, 1, 1), rs, std=math.sqrt(2.0 / ci))
    _bn_entries(sd, f"{p}.6", co, rs)


def sd_heads_state(sd: SD, rs, num_class: int) -> None:
    for name, chans in SD_HEADS:
        for j, (ci, co) in enumerate(zip(chans[:-1], chans[1:])):
            sepconv_state(sd, f"{name}.{j}", ci, co, rs)
    for fc in ("middle_fc1", "middle_fc2", "middle_fc3"):
        sd[fc + ".weight"] = torch.from_numpy((rs.standard_normal((num_class, 1280)) * 0.02).astype(np.float32))
        sd[fc + ".bias"] = torch.from_numpy((rs.standard_normal(num_class) * 0.02).astype(np.float32))


def build_sd_state(num_class: int = 83, temporal: str = "tsm", shift_div: int = 8, seed: int = 0) -> SD:
    sd = build_tsn_state(num_class, temporal, shift_div, seed)
    sd_heads_state(sd, np.random.RandomState(seed + 991), num_class)
    return sd


def sd_train_step(sd: SD, rgb, labels, num_segments=8, temporal="tsm", shift_div=8, bn_training=True,
                  alpha=0.1, beta=1e-6, temperature=3.0):
    """One reference-equivalent SD step (train_sd.py:217-282) on the oracle: returns (total, terms)."""
    for v in sd.values():
        if v.requires_grad:
            v.grad = None
    outs = sd_forward(rgb, sd, num_segments, temporal, shift_div, bn_training)
    total, terms = sd_loss(outs[:4], outs[4:], labels, alpha, beta, temperature)
    total.backward()
    return total.detach(), terms


# --------------------------------------------------------------------------------------------
# MTMM+SD combined wrapper — models/models_MTMM_SD.py:226-249 (ConvTranspose decoders), :431-532
# (forward), loss train_mtmm_sd.py:240-293.  ResNet-only in the reference; on MobileNetV2 the
# `maxpool` tap (64 ch @56^2) becomes the features[3] output (24 ch @56^2, the first SD tap) and
# `layer4` the features[18] output (1280 ch @7^2) — builder-defined like A10 / A13.
# --------------------------------------------------------------------------------------------
def convt_decoder(f, sd: SD, prefix: str, n_conv: int, bn_training: bool):
    """nn.Sequential(ConvTranspose2d(k4,s2,p1), BatchNorm2d, ..., ConvTranspose2d, Sigmoid): `n_conv`
    transposed convolutions (with bias), a BatchNorm after each but the last, NO activation in between
    (models/models_MTMM_SD.py:227-249)."""
    y = f
    for j in range(n_conv):
        y = F.conv_transpose2d(y, sd[f"{prefix}.{2 * j}.weight"], sd[f"{prefix}.{2 * j}.bias"], stride=2, padding=1)
        if j + 1 < n_conv:
            y = _bn(y, sd, f"{prefix}.{2 * j + 1}", bn_training)
    return torch.sigmoid(y)


def convt_decoder_state(sd: SD, prefix: str, chans, rs) -> None:
    n = len(chans) - 1
    for j, (ci, co) in enumerate(zip(chans[:-1], chans[1:])):
        sd[f"{prefix}.{2 * j}.weight"] = torch.from_numpy(
            (rs.standard_normal((ci, co, 4, 4)) * math.sqrt(1.0 / (4 * ci))).astype(np.float32))
        sd[f"{prefix}.{2 * j}.bias"] = torch.from_numpy((rs.standard_normal(co) * 0.1).astype(np.float32))
        if j + 1 < n:
            _bn_entries(sd, f"{prefix}.{2 * j + 1}", co, rs)


MTMM_SD_LOCAL = (24, 32, 1)            # local_decoder channels on MobileNetV2 (reference: 64, 32, 1)
MTMM_SD_GLOBAL = (1280, 256, 32, 1)    # global_decoder channels (reference: 2048, 256, 32, 1)


def build_mtmm_sd_state(num_class: int = 83, temporal: str = "tsm", shift_div: int = 8, seed: int = 0) -> SD:
    sd = build_sd_state(num_class, temporal, shift_div, seed)
    rs = np.random.RandomState(seed + 313)
    convt_decoder_state(sd, "local_decoder", MTMM_SD_LOCAL, rs)
    convt_decoder_state(sd, "global_decoder", MTMM_SD_GLOBAL, rs)
    return sd


def mtmm_sd_forward(x5, sd: SD, num_segments: int, temporal: str = "tsm", shift_div: int = 8, bn_training: bool = True):
    """-> the ten tensors of modal='rgb_depth' (models/models_MTMM_SD.py:522-523): output, mid1-3, final_fea,
    fea1-3, local_depth_out [NT,1,224,224], global_depth_out [NT,1,56,56]."""
    taps = {i: None for i in SD_TAPS + (18,)}
    x = x5.view((-1, 3) + tuple(x5.shape[-2:]))
    mobilenet_v2_features(x, sd, temporal, num_segments, shift_div, bn_training, taps=taps)
    mids, feas = [], []
    for (name, chans), tap, fc in zip(SD_HEADS, SD_TAPS, ("middle_fc1", "middle_fc2", "middle_fc3")):
        y = taps[tap]
        for j in range(len(chans) - 1):
            y = sepconv(y, sd, f"{name}.{j}", bn_training)
        fea = F.adaptive_avg_pool2d(y, 1)
        z = F.linear(torch.flatten(fea, 1), sd[fc + ".weight"], sd[fc + ".bias"])
        mids.append(z.view((-1, num_segments) + tuple(z.shape[1:])).mean(1))
        feas.append(fea)
    final_fea = F.adaptive_avg_pool2d(taps[18], 1)
    z = F.linear(torch.flatten(final_fea, 1), sd["new_fc.weight"], sd["new_fc.bias"])
    out = z.view((-1, num_segments) + tuple(z.shape[1:])).mean(1)
    local_depth = convt_decoder(taps[SD_TAPS[0]], sd, "local_decoder", len(MTMM_SD_LOCAL) - 1, bn_training)
    global_depth = convt_decoder(taps[18], sd, "global_decoder", len(MTMM_SD_GLOBAL) - 1, bn_training)
    return (out, *mids, final_fea, *feas, local_depth, global_depth)


def mtmm_sd_loss(outputs, feats, global_depth_out, depth_gt5, labels, alpha=0.1, beta=1e-6, temperature=3.0):
    """train_mtmm_sd.py:240-293: `loss` = CE(output) + 0.01 * MSE(g_depth_out, bilinear56(depth)); the SD
    terms as train_sd.py; total = (1-a)(loss + 3 CE) + a * 3 KD + b * 3 feature.  Returns (total, loss)."""
    out, m1, m2, m3 = outputs
    f4, f1, f2, f3 = feats
    gt = depth_gt5.view(-1, 1, depth_gt5.size(-2), depth_gt5.size(-1))
    gt = F.interpolate(gt, size=(56, 56), mode="bilinear")
    loss = F.cross_entropy(out, labels) + 0.01 * F.mse_loss(global_depth_out, gt)
    ce = [F.cross_entropy(o, labels) for o in (m1, m2, m3)]
    temp4 = torch.softmax(out / temperature, dim=1).detach()
    kd = [kd_loss(m, temp4, temperature) * temperature ** 2 for m in (m1, m2, m3)]
    fl = [feature_loss(f, f4.detach()) for f in (f1, f2, f3)]
    total = (1 - alpha) * (loss + sum(ce)) + alpha * sum(kd) + beta * sum(fl)
    return total, loss


def mtmm_sd_train_step(sd: SD, rgb, depth, labels, num_segments=8, temporal="tsm", shift_div=8, bn_training=True,
                       alpha=0.1, beta=1e-6, temperature=3.0):
    for v in sd.values():
        if v.requires_grad:
            v.grad = None
    outs = mtmm_sd_forward(rgb, sd, num_segments, temporal, shift_div, bn_training)
    total, loss = mtmm_sd_loss(outs[:4], outs[4:8], outs[9], depth, labels, alpha, beta, temperature)
    total.backward()           # the reference script calls loss.backward() (train_mtmm_sd.py:310, a bug noted in
    return total.detach(), loss.detach(), outs   # SURVEY section 4); the quantity it logs and means to train is total_loss


# --------------------------------------------------------------------------------------------
# EMA of the model state — train_mtmm.py:110-128 (EMAWrapper._update / update), called every step (:245)
# --------------------------------------------------------------------------------------------
def ema_update(ema_sd: SD, model_sd: SD, decay: float = 0.9999) -> None:
    """In place over EVERY state_dict entry, buffers and integer counters included: the reference computes
    ``decay * e + (1 - decay) * m`` with python-float scalars (so in the tensor's dtype for floating entries and in
    the default float32 for the int64 ``num_batches_tracked``) and ``copy_``s the result back (truncation)."""
    with torch.no_grad():
        for k, e in ema_sd.items():
            m = model_sd[k]
            e.copy_(decay * e + (1. - decay) * m)


# --------------------------------------------------------------------------------------------
# TemporalPool — models/temporal_shift.py:89-98: max over frames {2t-1, 2t, 2t+1} (clipped), stride 2
# --------------------------------------------------------------------------------------------
def temporal_pool_np(x: np.ndarray, n_segment: int) -> np.ndarray:
    nt, c, h, w = x.shape
    v = x.reshape(nt // n_segment, n_segment, c, h, w)
    t_out = (n_segment + 2 - 3) // 2 + 1
    out = np.empty((v.shape[0], t_out, c, h, w), x.dtype)
    for t in range(t_out):
        lo, hi = max(2 * t - 1, 0), min(2 * t + 1, n_segment - 1)
        out[:, t] = v[:, lo:hi + 1].max(axis=1)
    return out.reshape(-1, c, h, w)


# --------------------------------------------------------------------------------------------
# N3 — ResNet-50 bottleneck backbone.  The reference has no ResNet source of its own: models/models.py:108-117 builds
# `torchvision.models.resnet50` (third-party dependency, not vendored; torchvision 0.26.0 in this image, the v1.5
# Bottleneck with the stride on the 3x3 convolution: torchvision/models/resnet.py Bottleneck.forward / ResNet._forward_impl)
# and models/temporal_shift.py:101-146 (place='blockres') wraps `conv1` of every bottleneck with TemporalShift, which
# renames that weight to `...conv1.net.weight`.  Restated below over a state_dict with those key names; pinned against
# the live torchvision module + the reference's make_temporal_shift by tests/golden/make_golden.py -> resnet.npz.
# --------------------------------------------------------------------------------------------
RESNET50_LAYERS = (3, 4, 6, 3)


def resnet_bottleneck(x, sd: SD, prefix: str, stride: int, temporal: str, n_segment: int, shift_div: int, bn_training: bool):
    """torchvision Bottleneck.forward: relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1(x)))))))) + identity); conv1's input goes
    through the temporal shift when `temporal == 'tsm'` (TemporalShift.forward, models/temporal_shift.py:22-25)."""
    if temporal == "tsm":
        out = F.conv2d(temporal_shift(x, n_segment, shift_div), sd[prefix + ".conv1.net.weight"])
    else:
        out = F.conv2d(x, sd[prefix + ".conv1.weight"])
    out = F.relu(_bn(out, sd, prefix + ".bn1", bn_training))
    out = F.relu(_bn(F.conv2d(out, sd[prefix + ".conv2.weight"], stride=stride, padding=1), sd, prefix + ".bn2", bn_training))
    out = _bn(F.conv2d(out, sd[prefix + ".conv3.weight"]), sd, prefix + ".bn3", bn_training)
    identity = x
    if prefix + ".downsample.0.weight" in sd:
        identity = _bn(F.conv2d(x, sd[prefix + ".downsample.0.weight"], stride=stride), sd, prefix + ".downsample.1", bn_training)
    return F.relu(out + identity)


def resnet_features(x, sd: SD, layers=RESNET50_LAYERS, temporal: str = "tsm", n_segment: int = 8, shift_div: int = 8,
                    bn_training: bool = True, prefix: str = "base_model", with_stem: bool = False):
    """ResNet._forward_impl up to layer4: conv1 7x7/2 -> bn1 -> relu -> maxpool 3x3/2 -> layer1..4.  Returns the four stage
    outputs (the taps of models/models_SD.py:364-431; layer4 is the map models/models_MTMM.py:70-77 decodes), preceded by
    the max-pool output when `with_stem` (the tensor models/models_MTMM_SD.py:431-476 hands to local_decoder)."""
    y = F.relu(_bn(F.conv2d(x, sd[prefix + ".conv1.weight"], stride=2, padding=3), sd, prefix + ".bn1", bn_training))
    y = F.max_pool2d(y, kernel_size=3, stride=2, padding=1)
    outs = [y] if with_stem else []
    for si, n_blocks in enumerate(layers, 1):
        for bi in range(n_blocks):
            stride = 2 if (bi == 0 and si > 1) else 1
            y = resnet_bottleneck(y, sd, f"{prefix}.layer{si}.{bi}", stride, temporal, n_segment, shift_div, bn_training)
        outs.append(y)
    return tuple(outs)


def resnet_tsn_forward(x5, sd: SD, num_segments: int, layers=RESNET50_LAYERS, temporal: str = "tsm", shift_div: int = 8,
                       bn_training: bool = True):
    """models/models.py:323-356 with a ResNet base: fold T into the batch, backbone, AdaptiveAvgPool2d(1) + flatten, Dropout
    as identity, new_fc, mean over the segments."""
    x = x5.view((-1, 3) + tuple(x5.shape[-2:]))
    f = resnet_features(x, sd, layers, temporal, num_segments, shift_div, bn_training)[-1]
    z = F.linear(f.mean((2, 3)), sd["new_fc.weight"], sd["new_fc.bias"])
    return z.view((-1, num_segments) + tuple(z.shape[1:])).mean(dim=1)


def build_resnet_state(layers=RESNET50_LAYERS, num_class: int = 83, temporal: str = "tsm", seed: int = 0) -> SD:
    """state_dict of TSN(base_model='resnet50', dropout>0) with torchvision's key names (conv1 of every bottleneck under
    `.net` when a temporal module wraps it), loadable with strict=True.  Values from numpy RandomState(seed); the last
    BatchNorm of a block gets a small scale (as zero_init_residual would, but not zero) so that deep stacks stay O(1)."""
    rs = np.random.RandomState(seed)
    sd: SD = {}
    b = "base_model"
    _conv_entry(sd, f"{b}.conv1.weight", (64, 3, 7, 7), rs)
    _bn_entries(sd, f"{b}.bn1", 64, rs)
    inplanes = 64
    for si, n_blocks in enumerate(layers, 1):
        planes = 64 << (si - 1)
        for bi in range(n_blocks):
            p = f"{b}.layer{si}.{bi}"
            stride = 2 if (bi == 0 and si > 1) else 1
            _conv_entry(sd, p + (".conv1.net.weight" if temporal == "tsm" else ".conv1.weight"), (planes, inplanes, 1, 1), rs)
            _bn_entries(sd, p + ".bn1", planes, rs)
            _conv_entry(sd, p + ".conv2.weight", (planes, planes, 3, 3), rs)
            _bn_entries(sd, p + ".bn2", planes, rs)
            _conv_entry(sd, p + ".conv3.weight", (4 * planes, planes, 1, 1), rs)
            _bn_entries(sd, p + ".bn3", 4 * planes, rs)
            sd[p + ".bn3.weight"] *= 0.5
            if stride != 1 or inplanes != 4 * planes:
                _conv_entry(sd, p + ".downsample.0.weight", (4 * planes, inplanes, 1, 1), rs)
                _bn_entries(sd, p + ".downsample.1", 4 * planes, rs)
            inplanes = 4 * planes
    sd["new_fc.weight"] = torch.from_numpy((rs.standard_normal((num_class, inplanes)) * 0.02).astype(np.float32))
    sd["new_fc.bias"] = torch.from_numpy((rs.standard_normal(num_class) * 0.02).astype(np.float32))
    return sd


def build_resnet_mtmm_state(num_class: int = 83, temporal: str = "tsm", seed: int = 0, layers=RESNET50_LAYERS) -> SD:
    """models_MTMM.TSN(base_model='resnet50'): the backbone + new_fc + global_decoder over 2048 channels
    (models/models_MTMM.py:129-155 — the reference's own configuration of the decoder)."""
    sd = build_resnet_state(layers, num_class, temporal, seed)
    decoder_state(sd, np.random.RandomState(seed + 77), 512 * 4)
    return sd


def resnet_mtmm_train_step(sd: SD, rgb, depth, labels, num_segments=16, temporal="tsm", shift_div=8, bn_training=True,
                           layers=RESNET50_LAYERS):
    """One MTMM step (train_mtmm.py:205-245) on the ResNet-50 wrapper (models/models_MTMM.py:268-292): logits from the pooled
    layer4 map, depth from global_decoder(layer4), loss train_mtmm.py:223-231, backward.  Gradients are left in sd[*].grad."""
    for v in sd.values():
        if v.requires_grad:
            v.grad = None
    x = rgb.view((-1, 3) + tuple(rgb.shape[-2:]))
    f = resnet_features(x, sd, layers, temporal, num_segments, shift_div, bn_training)[-1]
    z = F.linear(f.mean((2, 3)), sd["new_fc.weight"], sd["new_fc.bias"])
    logits = z.view((-1, num_segments) + tuple(z.shape[1:])).mean(dim=1)
    dpred = global_decoder(f, sd, bn_training)
    loss, _ = mtmm_loss(logits, labels, dpred, depth)
    loss.backward()
    return loss.detach(), logits.detach(), dpred.detach()


# N3: the reference's OWN MTMM / SD wrappers are written for ResNet bases (models/models_MTMM.py:112-157, 268-292;
# models/models_SD.py:214-253, 364-431).  Restated over the ResNet functions above; pinned against the live, unmodified
# wrappers (is_shift=False: the reference's ResNet temporal module is Action) by make_golden.golden_resnet_wrappers.
RESNET_SD_HEADS = (("scala1", (256, 512, 1024, 2048)), ("scala2", (512, 1024, 2048)), ("scala3", (1024, 2048)))


def resnet_mtmm_forward(x5, sd: SD, num_segments: int, temporal: str = "tsm", shift_div: int = 8, bn_training: bool = True,
                        layers=RESNET50_LAYERS):
    """models_MTMM.TSN.forward: -> (logits [N,cls], depth [NT,1,8h,8w]) with h x w the layer4 map."""
    x = x5.view((-1, 3) + tuple(x5.shape[-2:]))
    f = resnet_features(x, sd, layers, temporal, num_segments, shift_div, bn_training)[-1]
    z = F.linear(f.mean((2, 3)), sd["new_fc.weight"], sd["new_fc.bias"])
    return z.view((-1, num_segments) + tuple(z.shape[1:])).mean(dim=1), global_decoder(f, sd, bn_training)


def resnet_sd_forward(x5, sd: SD, num_segments: int, temporal: str = "tsm", shift_div: int = 8, bn_training: bool = True,
                      layers=RESNET50_LAYERS, taps=None):
    """models_SD.TSN.forward: -> (output, mid1, mid2, mid3, final_fea, fea1, fea2, fea3), taps after layer1 / 2 / 3."""
    if taps is None:
        x = x5.view((-1, 3) + tuple(x5.shape[-2:]))
        taps = resnet_features(x, sd, layers, temporal, num_segments, shift_div, bn_training)
    t1, t2, t3, t4 = taps
    mids, feas = [], []
    for (name, chans), y, fc in zip(RESNET_SD_HEADS, (t1, t2, t3), ("middle_fc1", "middle_fc2", "middle_fc3")):
        for j in range(len(chans) - 1):
            y = sepconv(y, sd, f"{name}.{j}", bn_training)
        fea = F.adaptive_avg_pool2d(y, 1)
        z = F.linear(torch.flatten(fea, 1), sd[fc + ".weight"], sd[fc + ".bias"])
        mids.append(z.view((-1, num_segments) + tuple(z.shape[1:])).mean(1))
        feas.append(fea)
    final_fea = F.adaptive_avg_pool2d(t4, 1)
    z = F.linear(torch.flatten(final_fea, 1), sd["new_fc.weight"], sd["new_fc.bias"])
    return (z.view((-1, num_segments) + tuple(z.shape[1:])).mean(1), *mids, final_fea, *feas)


def build_resnet_sd_state(num_class: int = 83, temporal: str = "tsm", seed: int = 0, layers=RESNET50_LAYERS) -> SD:
    sd = build_resnet_state(layers, num_class, temporal, seed)
    rs = np.random.RandomState(seed + 991)
    for name, chans in RESNET_SD_HEADS:
        for j, (ci, co) in enumerate(zip(chans[:-1], chans[1:])):
            sepconv_state(sd, f"{name}.{j}", ci, co, rs)
    for fc in ("middle_fc1", "middle_fc2", "middle_fc3"):
        sd[fc + ".weight"] = torch.from_numpy((rs.standard_normal((num_class, 2048)) * 0.02).astype(np.float32))
        sd[fc + ".bias"] = torch.from_numpy((rs.standard_normal(num_class) * 0.02).astype(np.float32))
    return sd


def build_resnet_mtmm_sd_state(num_class: int = 83, temporal: str = "tsm", seed: int = 0) -> SD:
    """models_MTMM_SD.TSN(base_model='resnet50', modal='rgb_depth'): the SD network + local_decoder (64, 32, 1) on the
    max-pool output + global_decoder (2048, 256, 32, 1) on layer4 (models/models_MTMM_SD.py:226-249)."""
    sd = build_resnet_sd_state(num_class, temporal, seed)
    rs = np.random.RandomState(seed + 313)
    convt_decoder_state(sd, "local_decoder", (64, 32, 1), rs)
    convt_decoder_state(sd, "global_decoder", (2048, 256, 32, 1), rs)
    return sd


def resnet_mtmm_sd_forward(x5, sd: SD, num_segments: int, temporal: str = "tsm", shift_div: int = 8, bn_training: bool = True):
    """The ten tensors of models_MTMM_SD.TSN.forward (:522-523) from ONE backbone pass (the reference runs the backbone
    twice, by hand :431-476 and through a torch.fx extractor :492; both passes compute the same activations, only the
    BatchNorm running statistics are updated twice there)."""
    x = x5.view((-1, 3) + tuple(x5.shape[-2:]))
    pooled, t1, t2, t3, t4 = resnet_features(x, sd, RESNET50_LAYERS, temporal, num_segments, shift_div, bn_training, with_stem=True)
    outs = resnet_sd_forward(x5, sd, num_segments, temporal, shift_div, bn_training, taps=(t1, t2, t3, t4))
    return (*outs, convt_decoder(pooled, sd, "local_decoder", 2, bn_training),
            convt_decoder(t4, sd, "global_decoder", 3, bn_training))
