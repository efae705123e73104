"""TEST INFRASTRUCTURE — a numpy restatement of the C-ABI contracts in include/ehgr_b200.h, fp32 only, on HOST pointers.

Purpose: the host-side orchestration of the fused autograd Functions (which kernels are called, with which operands,
shapes and saved tensors, and where every gradient lands) can be exercised on a machine without a GPU: the CPU tests
monkeypatch ``_lib.call`` with ``call`` below and compare the result with PyTorch autograd on the same modules.  Nothing
in the product imports this file; the product path still refuses CPU tensors (``_lib.require_cuda``).  Each function
states the header contract it restates; the CUDA kernels themselves are checked against PyTorch in the ``-m gpu`` tests.
"""
import ctypes

import numpy as np

F32 = np.float32


def arr(ptr, shape, dtype=F32):
    """numpy view of host memory at `ptr` (an int address, or None / 0)."""
    if hasattr(ptr, "value"):
        ptr = ptr.value
    if not ptr:
        return None
    n = int(np.prod(shape))
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _struct(ref):
    return ref._obj if hasattr(ref, "_obj") else ref


def _act(z, code):
    if code == 1:
        return np.clip(z, 0.0, 6.0)
    if code == 2:
        return np.maximum(z, 0.0)
    return z


def _mask(z, code):
    if code == 1:
        return (z > 0) & (z < 6)
    if code == 2:
        return z > 0
    return np.ones_like(z, dtype=bool)


def rowop(ref, M, C):
    """struct ehgr_rowop evaluated for all rows: [M, C] float32 (C = K of the GEMM for CONV3)."""
    op = _struct(ref)
    mode = op.mode
    if mode == 0:
        return arr(op.in1, (M, C)).copy()
    if mode == 1:
        return _act(arr(op.in1, (M, C)) * arr(op.scale, (C,)) + arr(op.shift, (C,)), op.relu6).astype(F32)
    if mode == 2:
        T, hw, fold = op.n_segment, op.hw, op.fold
        x = arr(op.in1, (M // (T * hw), T, hw, C))
        out = np.zeros_like(x)
        if op.shift_dir >= 0:
            out[:, :-1, :, :fold] = x[:, 1:, :, :fold]
            out[:, 1:, :, fold:2 * fold] = x[:, :-1, :, fold:2 * fold]
        else:
            out[:, 1:, :, :fold] = x[:, :-1, :, :fold]
            out[:, :-1, :, fold:2 * fold] = x[:, 1:, :, fold:2 * fold]
        out[:, :, :, 2 * fold:] = x[:, :, :, 2 * fold:]
        return out.reshape(M, C)
    if mode == 3:
        g, r = arr(op.in1, (M, C)), arr(op.in2, (M, C))
        if op.relu6:
            g = g * _mask(r * arr(op.scale, (C,)) + arr(op.shift, (C,)), op.relu6)
        return (arr(op.ca, (C,)) * g + arr(op.cb, (C,)) * r + arr(op.cc, (C,))).astype(F32)
    if mode == 4:                       # GATE: in1 * (3 + g1[m] + g2[m / hw, c] + g3[m / hw, c])
        fr = M // op.hw
        x = arr(op.in1, (fr, op.hw, C))
        g = 3.0 + arr(op.in2, (fr, op.hw, 1)) + arr(op.scale, (fr, 1, C)) + arr(op.shift, (fr, 1, C))
        return (x * g).reshape(M, C).astype(F32)
    if mode == 5:
        H, W, cin, up = op.cv_h, op.cv_w, op.cv_cin, op.cv_up
        assert C == 9 * cin and op.hw == H * W
        frames = M // (H * W)
        src = arr(op.in1, (frames, H >> up, W >> up, cin))
        if op.scale:
            src = _act(src * arr(op.scale, (cin,)) + arr(op.shift, (cin,)), op.relu6)
        if up:
            src = src.repeat(2, axis=1).repeat(2, axis=2)
        pad = np.zeros((frames, H + 2, W + 2, cin), F32)
        pad[:, 1:-1, 1:-1] = src
        cols = [pad[:, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)]     # tap = 3*ky + kx
        return np.concatenate(cols, axis=-1).reshape(M, 9 * cin).astype(F32)
    raise NotImplementedError(f"row operand mode {mode}")


def _add_stats(stats, y, C):
    if stats:
        s = arr(stats, (2 * C,), np.float64)
        s[:C] += y.astype(np.float64).sum(0)
        s[C:] += (y.astype(np.float64) ** 2).sum(0)


def _half(v):
    return (v - 1) // 2 + 1


# ---- GEMM family -----------------------------------------------------------------------------------------------------
def ehgr_pw_gemm_bn(a, w, w16, w_is_kn, out, addend, stats, M, K, N, dtype, engine, fin, stream):
    assert dtype == 0 and fin is None
    A = rowop(a, M, K)
    Wm = arr(w, (K, N)) if w_is_kn else arr(w, (N, K)).T
    y = A @ Wm
    if addend:
        y = y + arr(addend, (M, N))
    arr(out, (M, N))[...] = y
    _add_stats(stats, y, N)


def ehgr_pw_gemm_w16(a, w, w16, w_is_kn, out, addend, stats, M, K, N, dtype, engine, stream):
    ehgr_pw_gemm_bn(a, w, w16, w_is_kn, out, addend, stats, M, K, N, dtype, engine, None, stream)


def ehgr_pw_wgrad(dy, a, dw, M, K, N, dtype, engine, stream):
    assert dtype == 0
    arr(dw, (N, K))[...] += rowop(dy, M, N).T @ rowop(a, M, K)


def ehgr_conv3_pack(w, wf, wd, cout, cin, dtype, stream):
    assert dtype == 0
    W = arr(w, (cout, cin, 9))
    if wf:
        arr(wf, (cout, 9, cin))[...] = W.transpose(0, 2, 1)
    if wd:
        arr(wd, (cin, 9, cout))[...] = W[:, :, ::-1].transpose(1, 2, 0)


def ehgr_conv3_unpack_grad(dwp, dw, cout, cin, stream):
    arr(dw, (cout, cin, 9))[...] += arr(dwp, (cout, 9, cin)).transpose(0, 2, 1)


# ---- BatchNorm -------------------------------------------------------------------------------------------------------
def ehgr_bn_finalize(stats, count, gamma, beta, rm, rv, momentum, eps, training, scale, shift, mean, invstd, c, stream):
    g = arr(gamma, (c,)) if gamma else np.ones(c, F32)
    b = arr(beta, (c,)) if beta else np.zeros(c, F32)
    if training:
        s = arr(stats, (2 * c,), np.float64)
        mu = s[:c] / count
        var = np.maximum(s[c:] / count - mu * mu, 0.0)
        if rm:
            arr(rm, (c,))[...] = (1 - momentum) * arr(rm, (c,)) + momentum * mu
            arr(rv, (c,))[...] = (1 - momentum) * arr(rv, (c,)) + momentum * var * (count / max(count - 1, 1))
    else:
        mu, var = arr(rm, (c,)).astype(np.float64), arr(rv, (c,)).astype(np.float64)
    istd = 1.0 / np.sqrt(var + eps)
    sc = g * istd
    arr(scale, (c,))[...] = sc
    arr(shift, (c,))[...] = b - mu * sc
    if mean:
        arr(mean, (c,))[...] = mu
    if invstd:
        arr(invstd, (c,))[...] = istd


def ehgr_bn_bwd_reduce_fin(g, raw, scale, shift, relu6, sums, m, c, dtype, fin, stream):
    assert dtype == 0 and fin is None
    G, R = arr(g, (m, c)).astype(np.float64), arr(raw, (m, c)).astype(np.float64)
    if relu6:
        G = G * _mask(R * arr(scale, (c,)) + arr(shift, (c,)), relu6)
    s = arr(sums, (2 * c,), np.float64)
    s[:c] += G.sum(0)
    s[c:] += (G * R).sum(0)


def ehgr_bn_bwd_finalize(sums, count, gamma, mean, invstd, training, ca, cb, cc, dgamma, dbeta, c, stream):
    s = arr(sums, (2 * c,), np.float64)
    sdz, sdzr = s[:c], s[c:]
    mu, istd = arr(mean, (c,)).astype(np.float64), arr(invstd, (c,)).astype(np.float64)
    g = arr(gamma, (c,)).astype(np.float64) if gamma else np.ones(c)
    sdzx = (sdzr - mu * sdz) * istd
    if dgamma:
        arr(dgamma, (c,))[...] = sdzx
    if dbeta:
        arr(dbeta, (c,))[...] = sdz
    sc = g * istd
    arr(ca, (c,))[...] = sc
    if training:
        k1, k2 = sdz / count, sdzx / count
        arr(cb, (c,))[...] = -sc * k2 * istd
        arr(cc, (c,))[...] = -sc * (k1 - mu * istd * k2)
    else:
        arr(cb, (c,))[...] = 0
        arr(cc, (c,))[...] = 0


def ehgr_row_apply(a, addend, out, m, c, dtype, stream):
    assert dtype == 0
    y = rowop(a, m, c)
    if addend:
        y = y + arr(addend, (m, c))
    arr(out, (m, c))[...] = y


# ---- N3: ResNet pieces (csrc/resnet.cu) ----------------------------------------------------------------------------------
def _patches(xp, k, stride, ho, wo):
    """xp [F, C, Hp, Wp] padded -> [F, C, k*k, ho, wo], tap = k*ky + kx."""
    return np.stack([xp[:, :, ky:ky + stride * (ho - 1) + 1:stride, kx:kx + stride * (wo - 1) + 1:stride]
                     for ky in range(k) for kx in range(k)], axis=2)


def ehgr_stem7_fwd(x, w, out, stats, frames, h, w_in, cout, x_dtype, out_dtype, stream):
    assert x_dtype == 0 and out_dtype == 0 and cout == 64
    ho, wo = _half(h), _half(w_in)
    xp = np.zeros((frames, 3, h + 6, w_in + 6), F32)
    xp[:, :, 3:-3, 3:-3] = arr(x, (frames, 3, h, w_in))
    P = _patches(xp, 7, 2, ho, wo)                                     # [F, 3, 49, ho, wo]
    y = np.einsum("fcthw,oct->fhwo", P, arr(w, (cout, 3, 49)), optimize=True).astype(F32)
    arr(out, (frames, ho, wo, cout))[...] = y
    _add_stats(stats, y.reshape(-1, cout), cout)


def ehgr_stem7_wgrad(dy, x, dw, frames, h, w_in, cout, x_dtype, dtype, stream):
    assert x_dtype == 0 and dtype == 0 and cout == 64
    ho, wo = _half(h), _half(w_in)
    D = rowop(dy, frames * ho * wo, cout).reshape(frames, ho, wo, cout)
    xp = np.zeros((frames, 3, h + 6, w_in + 6), F32)
    xp[:, :, 3:-3, 3:-3] = arr(x, (frames, 3, h, w_in))
    P = _patches(xp, 7, 2, ho, wo)
    arr(dw, (cout, 3, 49))[...] += np.einsum("fhwo,fcthw->oct", D, P, optimize=True)


def ehgr_stem7_im2col(x, a, frames, h, w_in, kp, x_dtype, dtype, stream):
    assert x_dtype == 0 and dtype == 0
    ho, wo = _half(h), _half(w_in)
    xp = np.zeros((frames, 3, h + 6, w_in + 6), F32)
    xp[:, :, 3:-3, 3:-3] = arr(x, (frames, 3, h, w_in))
    P = _patches(xp, 7, 2, ho, wo)                                     # [F, 3, 49, ho, wo]
    A = arr(a, (frames * ho * wo, kp))
    A[...] = 0
    A[:, :147] = P.transpose(0, 3, 4, 1, 2).reshape(frames * ho * wo, 147)


def ehgr_stem7_pack(w, wp, cout, kp, dtype, stream):
    assert dtype == 0
    Wp = arr(wp, (cout, kp))
    Wp[...] = 0
    Wp[:, :147] = arr(w, (cout, 147))


def ehgr_stem7_unpack_grad(dwp, dw, cout, kp, stream):
    arr(dw, (cout, 147))[...] += arr(dwp, (cout, kp))[:, :147]


def ehgr_maxpool3_fwd(a, y, idx, frames, h, w, c, dtype, stream):
    assert dtype == 0
    ho, wo = _half(h), _half(w)
    X = rowop(a, frames * h * w, c).reshape(frames, h, w, c).transpose(0, 3, 1, 2)
    xp = np.full((frames, c, h + 2, w + 2), -np.inf, F32)
    xp[:, :, 1:-1, 1:-1] = X
    P = _patches(xp, 3, 2, ho, wo)                                     # [F, C, 9, ho, wo]
    arr(y, (frames, ho, wo, c))[...] = P.max(2).transpose(0, 2, 3, 1)
    arr(idx, (frames, ho, wo, c), np.uint8)[...] = P.argmax(2).transpose(0, 2, 3, 1)   # first maximum in scan order


def ehgr_maxpool3_bwd(g, idx, gx, frames, h, w, c, dtype, stream):
    assert dtype == 0
    ho, wo = _half(h), _half(w)
    G, I = arr(g, (frames, ho, wo, c)), arr(idx, (frames, ho, wo, c), np.uint8)
    out = np.zeros((frames, h + 2, w + 2, c), F32)
    for tap in range(9):
        ky, kx = divmod(tap, 3)
        out[:, ky:ky + 2 * (ho - 1) + 1:2, kx:kx + 2 * (wo - 1) + 1:2] += G * (I == tap)
    arr(gx, (frames, h, w, c))[...] = out[:, 1:-1, 1:-1]


def ehgr_subsample2_fwd(x, y, stats, frames, h, w, c, dtype, stream):
    assert dtype == 0
    Y = arr(x, (frames, h, w, c))[:, ::2, ::2]
    arr(y, Y.shape)[...] = Y
    _add_stats(stats, Y.reshape(-1, c), c)


def ehgr_subsample2_bwd(g, gx, frames, h, w, c, dtype, stream):
    assert dtype == 0
    out = arr(gx, (frames, h, w, c))
    out[...] = 0
    out[:, ::2, ::2] = arr(g, (frames, _half(h), _half(w), c))


def ehgr_bn_add_relu(raw, scale, shift, addend, out, m, c, dtype, stream):
    assert dtype == 0
    arr(out, (m, c))[...] = np.maximum(arr(raw, (m, c)) * arr(scale, (c,)) + arr(shift, (c,)) + arr(addend, (m, c)), 0)


def ehgr_relu_bwd(g, out, gz, n, dtype, stream):
    assert dtype == 0
    arr(gz, (n,))[...] = arr(g, (n,)) * (arr(out, (n,)) > 0)


# ---- MobileNetV2 chain: depthwise (K7), stem (K9), head (K10), decoder pieces (N2) — restated with torch ops ------------------
def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a))


def ehgr_dw_fwd_bn(a, w, out, stats, nt, h, wd, c, stride, dtype, fin, stream):
    import torch.nn.functional as TF
    assert dtype == 0 and fin is None
    X = _t(rowop(a, nt * h * wd, c).reshape(nt, h, wd, c)).permute(0, 3, 1, 2)
    y = TF.conv2d(X, _t(arr(w, (c, 1, 3, 3))), stride=stride, padding=1, groups=c).permute(0, 2, 3, 1).numpy()
    arr(out, y.shape)[...] = y
    _add_stats(stats, y.reshape(-1, c), c)


def ehgr_dw_fwd(a, w, out, stats, nt, h, wd, c, stride, dtype, stream):
    ehgr_dw_fwd_bn(a, w, out, stats, nt, h, wd, c, stride, dtype, None, stream)


def ehgr_dw_bwd(dy, a, w, da, dw, nt, h, wd, c, stride, dtype, stream):
    import torch
    import torch.nn.functional as TF
    assert dtype == 0
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    X = _t(rowop(a, nt * h * wd, c).reshape(nt, h, wd, c)).permute(0, 3, 1, 2).requires_grad_(True)
    Wt = _t(arr(w, (c, 1, 3, 3)).copy()).requires_grad_(True)
    DY = _t(rowop(dy, nt * ho * wo, c).reshape(nt, ho, wo, c)).permute(0, 3, 1, 2)
    with torch.enable_grad():          # called from inside an autograd.Function.backward, where grad mode is off
        gx, gw = torch.autograd.grad(TF.conv2d(X, Wt, stride=stride, padding=1, groups=c), (X, Wt), DY)
    arr(da, (nt, h, wd, c))[...] = gx.permute(0, 2, 3, 1).numpy()
    arr(dw, (c, 1, 3, 3))[...] += gw.numpy()


def ehgr_stem_fwd_bn(x, w, out, stats, nt, h, wd, cout, x_dtype, dtype, fin, stream):
    import torch.nn.functional as TF
    assert x_dtype == 0 and dtype == 0 and fin is None
    y = TF.conv2d(_t(arr(x, (nt, 3, h, wd))), _t(arr(w, (cout, 3, 3, 3))), stride=2, padding=1).permute(0, 2, 3, 1).numpy()
    arr(out, y.shape)[...] = y
    _add_stats(stats, y.reshape(-1, cout), cout)


def ehgr_stem_wgrad(dy, x, dw, nt, h, wd, cout, x_dtype, dtype, stream):
    import torch
    import torch.nn.functional as TF
    assert x_dtype == 0 and dtype == 0
    ho, wo = (h - 1) // 2 + 1, (wd - 1) // 2 + 1
    Wt = torch.zeros(cout, 3, 3, 3, requires_grad=True)
    DY = _t(rowop(dy, nt * ho * wo, cout).reshape(nt, ho, wo, cout)).permute(0, 3, 1, 2)
    with torch.enable_grad():
        (gw,) = torch.autograd.grad(TF.conv2d(_t(arr(x, (nt, 3, h, wd))), Wt, stride=2, padding=1), (Wt,), DY)
    arr(dw, (cout, 3, 3, 3))[...] += gw.numpy()


def ehgr_pool_fwd(a, pooled, nt, hw, c, dtype, stream):
    assert dtype == 0
    arr(pooled, (nt, c))[...] = rowop(a, nt * hw, c).reshape(nt, hw, c).mean(1)


def ehgr_pool_bwd(dpooled, da, nt, hw, c, dtype, stream):
    assert dtype == 0
    arr(da, (nt, hw, c))[...] = arr(dpooled, (nt, 1, c)) / hw


def ehgr_fc_consensus_fwd(feat, w, bias, meanfeat, logits, n, T, f, k, stream):
    mf = arr(feat, (n, T, f)).mean(1)
    arr(meanfeat, (n, f))[...] = mf
    z = mf @ arr(w, (k, f)).T
    arr(logits, (n, k))[...] = z + (arr(bias, (k,)) if bias else 0)


def ehgr_fc_consensus_bwd(dlogits, meanfeat, w, dfeat, dw, dbias, n, T, f, k, stream):
    G = arr(dlogits, (n, k))
    arr(dfeat, (n, T, f))[...] = ((G @ arr(w, (k, f))) / T)[:, None, :]
    arr(dw, (k, f))[...] += G.T @ arr(meanfeat, (n, f))
    if dbias:
        arr(dbias, (k,))[...] += G.sum(0)


def ehgr_upsample2_fwd(x, y, frames, h, w, c, dtype, stream):
    assert dtype == 0
    arr(y, (frames, 2 * h, 2 * w, c))[...] = arr(x, (frames, h, w, c)).repeat(2, axis=1).repeat(2, axis=2)


def ehgr_upsample2_bwd(g_up, g, frames, h, w, c, dtype, stream):
    assert dtype == 0
    arr(g, (frames, h, w, c))[...] = arr(g_up, (frames, h, 2, w, 2, c)).sum((2, 4))


def ehgr_depth_head_fwd(a, w, bias, out, m, c, dtype, stream):
    assert dtype == 0
    z = rowop(a, m, c) @ arr(w, (c,)) + (arr(bias, (1,))[0] if bias else 0.0)
    arr(out, (m,))[...] = 1.0 / (1.0 + np.exp(-z))


def ehgr_depth_head_bwd(a, w, out, dout, g_a, dw, dbias, m, c, dtype, stream):
    assert dtype == 0
    o = arr(out, (m,))
    dz = arr(dout, (m,)) * o * (1 - o)
    arr(g_a, (m, c))[...] = dz[:, None] * arr(w, (c,))[None, :]
    arr(dw, (c,))[...] += dz @ rowop(a, m, c)
    if dbias:
        arr(dbias, (1,))[...] += dz.sum()


def ehgr_mtmm_loss(logits, labels, pred, depth_gt, depth_weight, loss_out, dlogits, dpred, n, k, frames, ph, pw, dtype, stream):
    """K12: total = CE + depth_weight * MSE(pred, mean of the 2x2 centre pixels of every 4x4 cell of depth_gt)."""
    import torch
    import torch.nn.functional as TF
    assert dtype == 0
    P = _t(arr(pred, (frames, ph, pw)).copy()).requires_grad_(True)
    gt = _t(arr(depth_gt, (frames, ph, 4, pw, 4)))[:, :, 1:3, :, 1:3].mean((2, 4))
    with torch.enable_grad():
        mse = ((P - gt) ** 2).mean()
        total = depth_weight * mse
        ce = torch.zeros(())
        if n > 0:
            L = _t(arr(logits, (n, k)).copy()).requires_grad_(True)
            ce = TF.cross_entropy(L, _t(arr(labels, (n,), np.int64)))
            total = total + ce
            gP, gL = torch.autograd.grad(total, (P, L))
            arr(dlogits, (n, k))[...] = gL.numpy()
        else:
            (gP,) = torch.autograd.grad(total, (P,))
    arr(dpred, (frames, ph, pw))[...] = gP.numpy()
    out = arr(loss_out, (3,))
    out[0] += float(total)
    out[1] += float(ce)
    out[2] += float(mse)


def ehgr_sd_loss(logits, feats, labels, alpha, beta, temperature, terms_out, dlogits, dfeats, n, k, rows, f, stream):
    """K13 (train_sd.py:178-193,227-265): total = (1-a) * sum_4 CE(z_i, y) + a * T^2 * sum_3 KD(z_i, softmax(z_0 / T).detach())
    + b * sum_3 sum((f_i - f_0.detach())^2 * [(f_i > 0) | (f_0 > 0)]);  terms = total, 4 CE, 3 KD (x T^2), 3 feature sums."""
    import torch
    import torch.nn.functional as TF
    y = _t(arr(labels, (n,), np.int64))
    Z = [_t(arr(p, (n, k)).copy()).requires_grad_(True) for p in logits]
    Fs = [_t(arr(p, (rows, f)).copy()).requires_grad_(True) for p in feats]
    with torch.enable_grad():
        ce = [TF.cross_entropy(z, y) for z in Z]
        soft = torch.softmax(Z[0] / temperature, dim=1).detach()
        kd = [-(torch.log_softmax(z / temperature, dim=1) * soft).sum(1).mean() * temperature ** 2 for z in Z[1:]]
        f0 = Fs[0].detach()
        fe = [(((fi - f0) ** 2) * ((fi > 0) | (f0 > 0)).float()).sum() for fi in Fs[1:]]
        total = (1 - alpha) * sum(ce) + alpha * sum(kd) + beta * sum(fe)
        grads = torch.autograd.grad(total, Z + Fs[1:])
    for p, g in zip(dlogits, grads[:4]):
        arr(p, (n, k))[...] = g.numpy()
    for p, g in zip(dfeats, grads[4:]):
        arr(p, (rows, f))[...] = g.numpy()
    arr(terms_out, (11,))[...] += np.array([float(total)] + [float(v) for v in ce + kd + fe], F32)


# ---- K2-K6: ACTION (models/action.py:61-116).  The host (action_ops.py) owns every buffer and only interprets qstats
# (-> ehgr_bn_finalize), bn3_sums (-> ehgr_bn_bwd_finalize), the three gates (GATE operand) and the gradients; the other
# workspaces are private to the kernels, so this restatement keeps its own (consistent) meaning for them: `dpool` is the
# gradient w.r.t. the spatial SUM `pool`, `dg1` stays sum_c gy*xs.
def _act_dims(a):
    return a.n, a.t, a.h, a.w, a.c, a.cr


def _act_params(a):
    c, cr = a.c, a.cr
    shapes = {"shift_w": (c, 1, 3), "p1_w": (1, 1, 3, 3, 3), "p2_squeeze": (cr, c), "p2_conv1": (cr, cr, 3), "p2_expand": (c, cr),
              "p3_squeeze": (cr, c), "p3_conv1": (cr, 1, 3, 3), "p3_expand": (c, cr)}
    return {k: _t(arr(getattr(a, k), shp).copy()) for k, shp in shapes.items()}


def _act_gates(a, mrow, pool, x3, P):
    """(g1 [M], g2 [nt, c], g3 [nt, c], s, u, pi) as differentiable torch functions of mrow [M], pool [nt, c] (spatial SUM),
    x3 [M, cr] (the BatchNorm output of the motion squeeze) and the parameters."""
    import torch
    import torch.nn.functional as TF
    n, t, h, w, c, cr = _act_dims(a)
    # spatio-temporal excitation (:76-83): sigmoid(conv3d over (t, h, w) of the channel mean)
    g1 = torch.sigmoid(TF.conv3d(mrow.view(n, 1, t, h, w), P["p1_w"], padding=1)).reshape(-1)
    # channel excitation (:86-96): spatial mean -> squeeze -> conv1d over t -> ReLU -> expand -> sigmoid
    sq = (pool / (h * w)).view(n, t, c) @ P["p2_squeeze"].t()                       # [n, t, cr]
    u = TF.conv1d(sq.transpose(1, 2), P["p2_conv1"], padding=1).transpose(1, 2)      # [n, t, cr]
    g2 = torch.sigmoid(torch.relu(u) @ P["p2_expand"].t()).reshape(n * t, c)
    # motion excitation (:99-113): depthwise 3x3 of frame t+1 minus frame t, zero for the last frame
    x3f = x3.view(n, t, h, w, cr).permute(0, 1, 4, 2, 3)                             # [n, t, cr, h, w]
    c3 = TF.conv2d(x3f.reshape(n * t, cr, h, w), P["p3_conv1"], padding=1, groups=cr).view(n, t, cr, h, w)
    d = torch.cat([c3[:, 1:] - x3f[:, :-1], torch.zeros_like(x3f[:, :1])], 1)
    pi = d.mean((3, 4))                                                              # [n, t, cr]
    g3 = torch.sigmoid(pi @ P["p3_expand"].t()).reshape(n * t, c)
    return g1, g2, g3, sq.reshape(n * t, cr), u.reshape(n * t, cr), pi.reshape(n * t, cr)


def ehgr_action_xs(a, x, xs, dtype, stream):
    assert dtype == 0
    a = _struct(a)
    n, t, h, w, c, cr = _act_dims(a)
    X = arr(x, (n, t, h * w, c))
    wk = arr(a.shift_w, (c, 3))
    pad = np.zeros((n, t + 2, h * w, c), F32)
    pad[:, 1:-1] = X
    Y = pad[:, :-2] * wk[:, 0] + pad[:, 1:-1] * wk[:, 1] + pad[:, 2:] * wk[:, 2]      # conv1d(k=3, padding=1) over t, per channel
    arr(xs, (n, t, h * w, c))[...] = Y
    arr(a.mrow, (n * t * h * w,))[...] = Y.mean(-1).reshape(-1)
    arr(a.pool, (n * t, c))[...] = Y.sum(2).reshape(n * t, c)
    q = Y.reshape(-1, c) @ arr(a.p3_squeeze, (cr, c)).T
    arr(a.q, (n * t * h * w, cr))[...] = q
    _add_stats(a.qstats, q, cr)


def ehgr_action_gates(a, stream):
    import torch
    a = _struct(a)
    n, t, h, w, c, cr = _act_dims(a)
    M = n * t * h * w
    x3 = _t(arr(a.q, (M, cr)) * arr(a.bn3_scale, (cr,)) + arr(a.bn3_shift, (cr,)))
    with torch.no_grad():
        g1, g2, g3, sq, u, pi = _act_gates(a, _t(arr(a.mrow, (M,)).copy()), _t(arr(a.pool, (n * t, c)).copy()), x3, _act_params(a))
    for name, v, shp in (("g1", g1, (M,)), ("g2", g2, (n * t, c)), ("g3", g3, (n * t, c)), ("s", sq, (n * t, cr)),
                         ("u", u, (n * t, cr)), ("pi", pi, (n * t, cr))):
        arr(getattr(a, name), shp)[...] = v.numpy()


def ehgr_action_bwd_reduce(a, gy, xs, dtype, stream):
    assert dtype == 0
    a = _struct(a)
    n, t, h, w, c, cr = _act_dims(a)
    prod = arr(gy, (n * t, h * w, c)) * arr(xs, (n * t, h * w, c))
    arr(a.dg1, (n * t * h * w,))[...] = prod.sum(-1).reshape(-1)
    arr(a.dgc, (n * t, c))[...] = prod.sum(1)


def _act_small_backward(a):
    """Gradients of sum(dg1*g1) + sum(dgc*(g2 + g3)) w.r.t. mrow, pool, x3 and the gate parameters."""
    import torch
    n, t, h, w, c, cr = _act_dims(a)
    M = n * t * h * w
    P = {k: v.requires_grad_(True) for k, v in _act_params(a).items()}
    mrow = _t(arr(a.mrow, (M,)).copy()).requires_grad_(True)
    pool = _t(arr(a.pool, (n * t, c)).copy()).requires_grad_(True)
    x3 = _t(arr(a.q, (M, cr)) * arr(a.bn3_scale, (cr,)) + arr(a.bn3_shift, (cr,))).requires_grad_(True)
    names = ("p1_w", "p2_squeeze", "p2_conv1", "p2_expand", "p3_conv1", "p3_expand")
    with torch.enable_grad():
        g1, g2, g3, _, _, _ = _act_gates(a, mrow, pool, x3, P)
        obj = (g1 * _t(arr(a.dg1, (M,)))).sum() + ((g2 + g3) * _t(arr(a.dgc, (n * t, c)))).sum()
        grads = torch.autograd.grad(obj, [mrow, pool, x3] + [P[k] for k in names])
    return grads[0], grads[1], grads[2], dict(zip(names, grads[3:]))


def ehgr_action_bwd_small(a, stream):
    a = _struct(a)
    n, t, h, w, c, cr = _act_dims(a)
    M = n * t * h * w
    dm, dpool, dx3, pg = _act_small_backward(a)
    arr(a.dm, (M,))[...] = dm.numpy()
    arr(a.dpool, (n * t, c))[...] = dpool.numpy()
    arr(a.dd, (n * t, cr))[...] = 0
    q = arr(a.q, (M, cr)).astype(np.float64)
    sums = arr(a.bn3_sums, (2 * cr,), np.float64)
    sums[:cr] += dx3.numpy().astype(np.float64).sum(0)
    sums[cr:] += (dx3.numpy().astype(np.float64) * q).sum(0)
    for k, g in pg.items():
        arr(getattr(a, "d_" + k), tuple(g.shape))[...] += g.numpy()


def ehgr_action_bwd_dxs(a, gy, xs, dxs, dtype, stream):
    assert dtype == 0
    a = _struct(a)
    n, t, h, w, c, cr = _act_dims(a)
    nt, hw = n * t, h * w
    M = nt * hw
    _, _, dx3, _ = _act_small_backward(a)
    q = arr(a.q, (M, cr))
    dq = arr(a.bn3_ca, (cr,)) * dx3.numpy() + arr(a.bn3_cb, (cr,)) * q + arr(a.bn3_cc, (cr,))       # BatchNorm backward
    XS = arr(xs, (M, c))
    G = 3.0 + arr(a.g1, (nt, hw, 1)) + arr(a.g2, (nt, 1, c)) + arr(a.g3, (nt, 1, c))
    out = arr(gy, (nt, hw, c)) * G + arr(a.dm, (nt, hw, 1)) / c + arr(a.dpool, (nt, 1, c))
    arr(dxs, (M, c))[...] = out.reshape(M, c) + dq @ arr(a.p3_squeeze, (cr, c))
    arr(a.d_p3_squeeze, (cr, c))[...] += dq.T @ XS


def ehgr_action_fir_bwd(a, dxs, x, addend, dx, dtype, stream):
    assert dtype == 0
    a = _struct(a)
    n, t, h, w, c, cr = _act_dims(a)
    D = arr(dxs, (n, t, h * w, c))
    X = arr(x, (n, t, h * w, c))
    wk = arr(a.shift_w, (c, 3))
    dp = np.zeros((n, t + 2, h * w, c), F32)
    dp[:, 1:-1] = D
    out = dp[:, 2:] * wk[:, 0] + dp[:, 1:-1] * wk[:, 1] + dp[:, :-2] * wk[:, 2]       # the adjoint of the temporal FIR
    if addend:
        out = out + arr(addend, (n, t, h * w, c))
    arr(dx, (n, t, h * w, c))[...] = out
    xp = np.zeros((n, t + 2, h * w, c), F32)
    xp[:, 1:-1] = X
    gw = arr(a.d_shift_w, (c, 3))
    for k in range(3):
        gw[:, k] += (D * xp[:, k:k + t]).sum((0, 1, 2))


# ---- the remaining entry points of include/ehgr_b200.h ------------------------------------------------------------------
def _shift(x, out, n_batch, T, c, hw, fold, dtype, layout, direction):
    assert dtype == 0
    if n_batch * T * c * hw == 0:
        return
    shape = (n_batch, T, c, hw) if layout == 0 else (n_batch, T, hw, c)
    X, Y = arr(x, shape), arr(out, shape)
    if layout == 1:
        X, Y = X.transpose(0, 1, 3, 2), Y.transpose(0, 1, 3, 2)          # views with the channel axis third
    Y[...] = 0
    a, b = (slice(None, -1), slice(1, None)) if direction > 0 else (slice(1, None), slice(None, -1))
    Y[:, a, :fold] = X[:, b, :fold]
    Y[:, b, fold:2 * fold] = X[:, a, fold:2 * fold]
    Y[:, :, 2 * fold:] = X[:, :, 2 * fold:]


def ehgr_temporal_shift_fwd(x, out, n_batch, n_segment, c, hw, fold, dtype, layout, stream):
    """K1: out[:, t, :fold] = x[:, t+1, :fold]; out[:, t, fold:2fold] = x[:, t-1, fold:2fold]; rest copied; zero at the ends."""
    _shift(x, out, n_batch, n_segment, c, hw, fold, dtype, layout, +1)


def ehgr_temporal_shift_bwd(g, gx, n_batch, n_segment, c, hw, fold, dtype, layout, stream):
    _shift(g, gx, n_batch, n_segment, c, hw, fold, dtype, layout, -1)


def ehgr_pw_gemm(a, w, w_is_kn, out, addend, stats, M, K, N, dtype, engine, stream):
    ehgr_pw_gemm_bn(a, w, None, w_is_kn, out, addend, stats, M, K, N, dtype, engine, None, stream)


def ehgr_stem_fwd(x, w, out, stats, nt, h, wd, cout, x_dtype, dtype, stream):
    ehgr_stem_fwd_bn(x, w, out, stats, nt, h, wd, cout, x_dtype, dtype, None, stream)


def ehgr_bn_bwd_reduce(g, raw, scale, shift, relu6, sums, m, c, dtype, stream):
    ehgr_bn_bwd_reduce_fin(g, raw, scale, shift, relu6, sums, m, c, dtype, None, stream)


def ehgr_dw_dgrad(dy, w, da, nt, h, wd, c, stride, dtype, stream):
    import torch
    import torch.nn.functional as TF
    assert dtype == 0
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    X = torch.zeros(nt, c, h, wd, requires_grad=True)
    DY = _t(rowop(dy, nt * ho * wo, c).reshape(nt, ho, wo, c)).permute(0, 3, 1, 2)
    with torch.enable_grad():
        (gx,) = torch.autograd.grad(TF.conv2d(X, _t(arr(w, (c, 1, 3, 3))), stride=stride, padding=1, groups=c), (X,), DY)
    arr(da, (nt, h, wd, c))[...] = gx.permute(0, 2, 3, 1).numpy()


def ehgr_dw_wgrad(dy, a, dw, nt, h, wd, c, stride, dtype, stream):
    import torch
    import torch.nn.functional as TF
    assert dtype == 0
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    X = _t(rowop(a, nt * h * wd, c).reshape(nt, h, wd, c)).permute(0, 3, 1, 2)
    Wt = torch.zeros(c, 1, 3, 3, requires_grad=True)
    DY = _t(rowop(dy, nt * ho * wo, c).reshape(nt, ho, wo, c)).permute(0, 3, 1, 2)
    with torch.enable_grad():
        (gw,) = torch.autograd.grad(TF.conv2d(X, Wt, stride=stride, padding=1, groups=c), (Wt,), DY)
    arr(dw, (c, 1, 3, 3))[...] += gw.numpy()


def ehgr_normalize_u8(src, dst, n_planes, channels, plane, mean, stdv, div, dst_dtype, stream):
    """N4: dst[p][i] = (float(src[p][i]) / div - mean[p % channels]) / std[p % channels], every step in fp32."""
    assert dst_dtype == 0
    v = arr(src, (n_planes, plane), np.uint8).astype(F32) / F32(div)
    ch = np.arange(n_planes) % channels
    if mean:
        v = v - arr(mean, (channels,))[ch][:, None]
    if stdv:
        v = v / arr(stdv, (channels,))[ch][:, None]
    arr(dst, (n_planes, plane))[...] = v.astype(F32)


def ehgr_temporal_pool_fwd(x, out, n, t_in, frame_elems, dtype, stream):
    """N4: max over the frames {2t'-1, 2t', 2t'+1} of a clip (max_pool3d (3,1,1) / (2,1,1) / (1,0,0))."""
    assert dtype == 0
    t_out = (t_in - 1) // 2 + 1
    X = arr(x, (n, t_in, frame_elems))
    pad = np.full((n, t_in + 2, frame_elems), -np.inf, F32)
    pad[:, 1:-1] = X
    arr(out, (n, t_out, frame_elems))[...] = np.stack([pad[:, 2 * k:2 * k + 3].max(1) for k in range(t_out)], 1)


def ehgr_temporal_pool_bwd(x, g, dx, n, t_in, frame_elems, dtype, stream):
    assert dtype == 0
    t_out = (t_in - 1) // 2 + 1
    X, G = arr(x, (n, t_in, frame_elems)), arr(g, (n, t_out, frame_elems))
    pad = np.full((n, t_in + 2, frame_elems), -np.inf, F32)
    pad[:, 1:-1] = X
    acc = np.zeros((n, t_in + 2, frame_elems), F32)
    for k in range(t_out):
        win = pad[:, 2 * k:2 * k + 3]
        first = win.argmax(1)                       # first maximum of the window (strict '>' scan)
        for j in range(3):
            acc[:, 2 * k + j] += G[:, k] * (first == j)
    arr(dx, (n, t_in, frame_elems))[...] = acc[:, 1:-1]


def ehgr_sgd_step(p, g, buf, code, lr_mult, decay_mult, n_groups, lr_dev, momentum, weight_decay, n, ema, ema_decay, p16, stream):
    """N1: d = g + wd * decay_mult[k] * p; buf = momentum * buf + d; p -= lr * lr_mult[k] * buf (fp32, torch.optim.SGD's
    order of operations); optional EMA of the stepped parameter (two fp32 products, one fp32 sum)."""
    assert not p16, "the bf16 mirror only exists on the GPU"
    P, G, B = arr(p, (n,)), arr(g, (n,)), arr(buf, (n,))
    k = arr(code, (n,), np.uint8)
    live = k != 255
    kk = np.where(live, k, 0).astype(np.int64)
    lr = arr(lr_dev, (1,))[0]
    d = (G + (F32(weight_decay) * arr(decay_mult, (n_groups,))[kk]) * P).astype(F32)
    nb = (F32(momentum) * B + d).astype(F32)
    npar = (P - (lr * arr(lr_mult, (n_groups,))[kk]).astype(F32) * nb).astype(F32)
    B[live] = nb[live]
    P[live] = npar[live]
    if ema:
        E = arr(ema, (n,))
        dec, rest = F32(ema_decay), F32(1.0 - float(ema_decay))      # python scalars reach fp32 as (float)decay, (float)(1. - decay)
        E[live] = (dec * E + rest * P).astype(F32)[live]


def ehgr_ema_update(ema, x, n, decay, is_int64, stream):
    dec, rest = F32(decay), F32(1.0 - float(decay))
    if is_int64:
        E, X = arr(ema, (n,), np.int64), arr(x, (n,), np.int64)
        E[...] = (dec * E.astype(F32) + rest * X.astype(F32)).astype(F32).astype(np.int64)
    else:
        E, X = arr(ema, (n,)), arr(x, (n,))
        E[...] = (dec * E + rest * X).astype(F32)


_TABLE = {k: v for k, v in globals().items() if k.startswith("ehgr_")}


def call(name, *args, algo_bytes=0, algo_flops=0, tag=""):
    """Drop-in for ehgr_b200._lib.call on host tensors."""
    if name not in _TABLE:
        raise NotImplementedError(f"the ABI emulator has no restatement of {name}")
    _TABLE[name](*args)
