"""Per-kernel parity: each C-ABI entry point against a plain PyTorch restatement of the same op
evaluated in fp64 on the CPU (fp32 kernels: 1e-5 relative; bf16 storage: 2e-2)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


def _E():
    import ehgr_b200
    return ehgr_b200


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def _rows(x_nchw):
    """[NT,C,H,W] -> [M, C] rows (NHWC order)."""
    return x_nchw.permute(0, 2, 3, 1).reshape(-1, x_nchw.shape[1])


def _dev_rows(t, dtype):
    return t.to(dtype).cuda().contiguous()


def _call(name, *args):
    _E()._lib.call(name, *args)


def _sp():
    return _E()._lib.stream_ptr(torch.device("cuda"))


# ---------------------------------------------------------------------------------------------
# row operand semantics, through ehgr_row_apply
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_row_apply_modes(dtype):
    E = _E()
    f = E.fused
    M, C, T, hw = 2 * 4 * 6, 24, 4, 6
    x = _rand((M, C), 1).to(dtype)
    xr = x.double()
    scale, shift = _rand((C,), 2).abs() + 0.5, _rand((C,), 3)
    add = _rand((M, C), 4).to(dtype)
    xd, ad = x.cuda(), add.cuda()
    sd_, sh_ = scale.cuda(), shift.cuda()
    code = E._lib.dtype_code(xd)

    def run(op, addend=None):
        out = torch.empty_like(xd)
        _call("ehgr_row_apply", ctypes.byref(op), 0 if addend is None else addend.data_ptr(), out.data_ptr(), M, C, code, _sp())
        return out.cpu().double()

    assert rel_err(run(f.op_plain(xd), ad), xr + add.double()) < TOL[dtype]
    want = torch.clamp(xr * scale.double() + shift.double(), 0, 6)
    assert rel_err(run(f.op_affine(xd, sd_, sh_, True)), want) < TOL[dtype]
    want = xr * scale.double() + shift.double() + add.double()
    assert rel_err(run(f.op_affine(xd, sd_, sh_, False), ad), want) < TOL[dtype]
    # shift (NHWC rows) == reference shift on the NCHW view, exactly
    from oracle import ref_oracle as O
    x4 = x.view(2 * T, 2, 3, C).permute(0, 3, 1, 2).contiguous()          # [NT, C, 2, 3]
    for fold in (3, 8, 0, 12):
        want = torch.from_numpy(O.temporal_shift_np(x4.float().numpy(), T, C // fold if fold else 10 ** 6)) if fold else x4.float()
        if fold:
            # oracle takes fold_div; emulate an absolute fold by picking c//fold_div == fold
            div = C // fold
            assert C // div == fold
            want = torch.from_numpy(O.temporal_shift_np(x4.float().numpy(), T, div))
        got = run(f.op_shift(xd, T, fold, hw, 1)).float().view(2 * T, 2, 3, C).permute(0, 3, 1, 2)
        assert torch.equal(got, want.float()), fold
        if fold:
            wantb = torch.from_numpy(O.temporal_shift_bwd_np(x4.float().numpy(), T, div))
            gotb = run(f.op_shift(xd, T, fold, hw, -1)).float().view(2 * T, 2, 3, C).permute(0, 3, 1, 2)
            assert torch.equal(gotb, wantb.float())
    # BNBWD
    raw = _rand((M, C), 5).to(dtype)
    ca, cb, cc = _rand((C,), 6), _rand((C,), 7) * 0.1, _rand((C,), 8) * 0.1
    z = raw.double() * scale.double() + shift.double()
    mask = ((z > 0) & (z < 6)).double()
    want = ca.double() * mask * xr + cb.double() * raw.double() + cc.double()
    dev = [t.cuda() for t in (raw, ca, cb, cc)]
    got = run(f.op_bnbwd(xd, *dev, sd_, sh_, True))
    assert rel_err(got, want) < TOL[dtype]


# ---------------------------------------------------------------------------------------------
# pointwise GEMM (SIMT engine) fwd / dgrad / wgrad
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,K,N", [(200, 16, 96), (333, 24, 144), (128, 144, 24), (50, 960, 320), (1000, 32, 192), (64, 320, 1280)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pw_gemm_simt(M, K, N, dtype):
    E = _E()
    f = E.fused
    a = _rand((M, K), 10).to(dtype)
    w = _rand((N, K), 11, (2.0 / K) ** 0.5)
    add = _rand((M, N), 12).to(dtype)
    scale, shift = _rand((K,), 13).abs() + 0.5, _rand((K,), 14)
    ad, wd, addd = a.cuda(), w.cuda(), add.cuda()
    sc_d, sh_d = scale.cuda(), shift.cuda()      # device tensors stay referenced until the kernels have run
    code = E._lib.dtype_code(ad)
    # forward with lazy BN+ReLU6 prologue, stats
    out = torch.empty((M, N), dtype=dtype, device="cuda")
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    _call("ehgr_pw_gemm", ctypes.byref(f.op_affine(ad, sc_d, sh_d, True)), wd.data_ptr(), 0, out.data_ptr(), 0,
          stats.data_ptr(), M, K, N, code, 1, _sp())
    aa = torch.clamp(a.double() * scale.double() + shift.double(), 0, 6)
    want = aa @ w.double().t()
    assert rel_err(out.cpu(), want) < TOL[dtype]
    assert rel_err(stats[:N].cpu(), want.sum(0)) < 1e-4 and rel_err(stats[N:].cpu(), (want ** 2).sum(0)) < 1e-4
    # "dgrad" form: out[M,K] = g[M,N] @ W[N,K] + addend
    g = _rand((M, N), 15).to(dtype)
    out2 = torch.empty((M, K), dtype=dtype, device="cuda")
    add2 = _rand((M, K), 16).to(dtype)
    gd, add2d = g.cuda(), add2.cuda()
    _call("ehgr_pw_gemm", ctypes.byref(f.op_plain(gd)), wd.data_ptr(), 1, out2.data_ptr(), add2d.data_ptr(), 0,
          M, N, K, code, 1, _sp())
    assert rel_err(out2.cpu(), g.double() @ w.double() + add2.double()) < TOL[dtype]
    # wgrad: dW[N,K] = g^T a'
    dw = torch.zeros((N, K), dtype=torch.float32, device="cuda")
    _call("ehgr_pw_wgrad", ctypes.byref(f.op_plain(gd)), ctypes.byref(f.op_affine(ad, sc_d, sh_d, True)),
          dw.data_ptr(), M, K, N, code, 1, _sp())
    assert rel_err(dw.cpu(), g.double().t() @ aa) < max(TOL[dtype], 2e-5)


# ---------------------------------------------------------------------------------------------
# pointwise GEMM, tcgen05 engine (bf16): every MobileNetV2 (K, N) pair class, ragged M, padded K / N
# ---------------------------------------------------------------------------------------------
TC_SHAPES = [(300, 16, 96), (1000, 24, 144), (257, 96, 24), (128, 144, 32), (5000, 32, 192), (130, 192, 64),
             (700, 64, 384), (260, 384, 96), (400, 576, 160), (140, 160, 960), (98, 960, 320), (392, 320, 1280),
             (64, 1280, 320), (20000, 32, 16), (1, 16, 16)]


@pytest.mark.parametrize("M,K,N", TC_SHAPES)
def test_pw_gemm_tcgen05_forward(M, K, N):
    E = _E()
    f = E.fused
    dtype = torch.bfloat16
    a = _rand((M, K), 60).to(dtype)
    w = _rand((N, K), 61, (2.0 / K) ** 0.5)
    scale, shift = _rand((K,), 62).abs() + 0.5, _rand((K,), 63)
    ad, wd, sc_d, sh_d = a.cuda(), w.cuda(), scale.cuda(), shift.cuda()
    out = torch.full((M, N), float("nan"), dtype=dtype, device="cuda")
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    _call("ehgr_pw_gemm", ctypes.byref(f.op_affine(ad, sc_d, sh_d, True)), wd.data_ptr(), 0, out.data_ptr(), 0,
          stats.data_ptr(), M, K, N, 1, 2, _sp())
    torch.cuda.synchronize()
    aa = torch.clamp(a.double() * scale.double() + shift.double(), 0, 6).to(dtype).double()   # operand is rounded to bf16
    want = aa @ w.to(dtype).double().t()
    assert rel_err(out.cpu(), want) < 1e-2
    # the batch statistics describe the STORED (bf16-rounded) tensor — the one the consumer normalises
    stored = out.cpu().double()
    assert rel_err(stats[:N].cpu(), stored.sum(0)) < 1e-5 and rel_err(stats[N:].cpu(), (stored ** 2).sum(0)) < 1e-5
    assert rel_err(stats[:N].cpu(), want.sum(0)) < 5e-3 and rel_err(stats[N:].cpu(), (want ** 2).sum(0)) < 5e-3
    # same call on the SIMT engine must agree to bf16 rounding
    out_s = torch.empty_like(out)
    _call("ehgr_pw_gemm", ctypes.byref(f.op_affine(ad, sc_d, sh_d, True)), wd.data_ptr(), 0, out_s.data_ptr(), 0, 0,
          M, K, N, 1, 1, _sp())
    assert rel_err(out.cpu(), out_s.cpu().double()) < 2e-2


@pytest.mark.parametrize("M,K,N", TC_SHAPES)
def test_pw_gemm_tcgen05_dgrad_form(M, K, N):
    """out[M,K] = rowop(g)[M,N] @ W[N,K] + addend: the conv weight read as an MN-major B operand."""
    E = _E()
    f = E.fused
    dtype = torch.bfloat16
    g = _rand((M, N), 64).to(dtype)
    raw = _rand((M, N), 65).to(dtype)
    w = _rand((N, K), 66, (2.0 / K) ** 0.5)
    add = _rand((M, K), 67).to(dtype)
    ca, cb, cc = _rand((N,), 68), _rand((N,), 69) * 0.1, _rand((N,), 70) * 0.1
    scale, shift = _rand((N,), 71).abs() + 0.5, _rand((N,), 72)
    out = torch.full((M, K), float("nan"), dtype=dtype, device="cuda")
    dev = [t.cuda() for t in (g, raw, ca, cb, cc, scale, shift, w, add)]
    op = f.op_bnbwd(*dev[:7], True)
    _call("ehgr_pw_gemm", ctypes.byref(op), dev[7].data_ptr(), 1, out.data_ptr(), dev[8].data_ptr(), 0, M, N, K, 1, 2, _sp())
    torch.cuda.synchronize()
    z = raw.double() * scale.double() + shift.double()
    mask = ((z > 0) & (z < 6)).double()
    dy = (ca.double() * mask * g.double() + cb.double() * raw.double() + cc.double()).to(dtype).double()
    want = dy @ w.to(dtype).double() + add.double()
    assert rel_err(out.cpu(), want) < 1e-2


@pytest.mark.parametrize("M,K,N", TC_SHAPES)
def test_pw_wgrad_tcgen05(M, K, N):
    """dW[N,K] = rowop(dy)^T rowop(a): both operands MN-major, accumulation over row tiles in TMEM."""
    E = _E()
    f = E.fused
    dtype = torch.bfloat16
    g = _rand((M, N), 80).to(dtype)
    raw = _rand((M, N), 81).to(dtype)
    a = _rand((M, K), 82).to(dtype)
    ca, cb, cc = _rand((N,), 83), _rand((N,), 84) * 0.1, _rand((N,), 85) * 0.1
    s_o, b_o = _rand((N,), 86).abs() + 0.5, _rand((N,), 87)
    s_a, b_a = _rand((K,), 88).abs() + 0.5, _rand((K,), 89)
    dev = [t.cuda() for t in (g, raw, ca, cb, cc, s_o, b_o, a, s_a, b_a)]
    dy_op = f.op_bnbwd(*dev[:7], True)
    a_op = f.op_affine(dev[7], dev[8], dev[9], True)
    dw = torch.zeros((N, K), dtype=torch.float32, device="cuda")
    _call("ehgr_pw_wgrad", ctypes.byref(dy_op), ctypes.byref(a_op), dw.data_ptr(), M, K, N, 1, 2, _sp())
    torch.cuda.synchronize()
    z = raw.double() * s_o.double() + b_o.double()
    mask = ((z > 0) & (z < 6)).double()
    dy = (ca.double() * mask * g.double() + cb.double() * raw.double() + cc.double()).to(dtype).double()
    aa = torch.clamp(a.double() * s_a.double() + b_a.double(), 0, 6).to(dtype).double()
    want = dy.t() @ aa
    assert rel_err(dw.cpu(), want) < 5e-3
    dw_s = torch.zeros_like(dw)
    _call("ehgr_pw_wgrad", ctypes.byref(dy_op), ctypes.byref(a_op), dw_s.data_ptr(), M, K, N, 1, 1, _sp())
    assert rel_err(dw.cpu(), dw_s.cpu().double()) < 2e-2


@pytest.mark.parametrize("K,N", [(16, 96), (24, 144), (64, 384), (160, 960), (96, 24), (576, 160)])
@pytest.mark.parametrize("a_mode", ["plain", "affine", "shift"])
def test_pw_wgrad_tcgen05_async_operands(K, N, a_mode):
    """The operand combinations the fused chain issues (dy materialised, a plain / lazily normalised / shifted):
    staged with cp.async straight into the MN-major core-matrix layout."""
    E = _E()
    f = E.fused
    dtype = torch.bfloat16
    T, hw, clips = 4, 49, 3
    M = clips * T * hw
    g = _rand((M, N), 90).to(dtype)
    a = _rand((M, K), 91).to(dtype)
    s_a, b_a = _rand((K,), 92).abs() + 0.5, _rand((K,), 93)
    gd, ad, sd_, bd = g.cuda(), a.cuda(), s_a.cuda(), b_a.cuda()
    if a_mode == "plain":
        a_op, aa = f.op_plain(ad), a.double()
    elif a_mode == "affine":
        a_op = f.op_affine(ad, sd_, bd, True)
        aa = torch.clamp(a.double() * s_a.double() + b_a.double(), 0, 6).to(dtype).double()
    else:
        fold = K // 8
        a_op = f.op_shift(ad, T, fold, hw)
        x5 = a.double().view(clips, T, hw, K)
        sh = torch.zeros_like(x5)
        sh[:, :-1, :, :fold] = x5[:, 1:, :, :fold]
        sh[:, 1:, :, fold:2 * fold] = x5[:, :-1, :, fold:2 * fold]
        sh[:, :, :, 2 * fold:] = x5[:, :, :, 2 * fold:]
        aa = sh.view(M, K)
    dw = torch.zeros((N, K), dtype=torch.float32, device="cuda")
    _call("ehgr_pw_wgrad", ctypes.byref(f.op_plain(gd)), ctypes.byref(a_op), dw.data_ptr(), M, K, N, 1, 2, _sp())
    torch.cuda.synchronize()
    assert rel_err(dw.cpu(), g.double().t() @ aa) < 5e-3


@pytest.mark.parametrize("K,N", [(16, 96), (24, 144), (32, 192), (64, 384), (160, 960), (96, 24)])
def test_pw_gemm_tcgen05_shift_operand(K, N):
    """SHIFT as a cp.async gather (incl. channel vectors that straddle a fold boundary), ragged last tile."""
    E = _E()
    f = E.fused
    dtype = torch.bfloat16
    T, hw, clips = 8, 49, 2
    M = clips * T * hw
    a = _rand((M, K), 94).to(dtype)
    w = _rand((N, K), 95, (2.0 / K) ** 0.5)
    ad, wd = a.cuda(), w.cuda()
    fold = K // 8
    x5 = a.double().view(clips, T, hw, K)
    sh = torch.zeros_like(x5)
    sh[:, :-1, :, :fold] = x5[:, 1:, :, :fold]
    sh[:, 1:, :, fold:2 * fold] = x5[:, :-1, :, fold:2 * fold]
    sh[:, :, :, 2 * fold:] = x5[:, :, :, 2 * fold:]
    out = torch.full((M, N), float("nan"), dtype=dtype, device="cuda")
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    _call("ehgr_pw_gemm", ctypes.byref(f.op_shift(ad, T, fold, hw)), wd.data_ptr(), 0, out.data_ptr(), 0, stats.data_ptr(),
          M, K, N, 1, 2, _sp())
    torch.cuda.synchronize()
    want = sh.view(M, K) @ w.to(dtype).double().t()
    assert rel_err(out.cpu(), want) < 1e-2
    assert rel_err(stats[:N].cpu(), want.sum(0)) < 2e-3


@pytest.mark.parametrize("M,K,N", [(300, 16, 96), (257, 96, 24), (400, 576, 160), (98, 960, 320), (392, 320, 1280), (64, 1280, 320)])
@pytest.mark.parametrize("kn", [0, 1])
def test_pw_gemm_w16_matches_fp32_weight_path(M, K, N, kn):
    """ehgr_pw_gemm_w16 (bf16 weight mirror staged with cp.async; resident and streamed B) gives bit-identical
    results to ehgr_pw_gemm (fp32 weights converted in the kernel): both round the weights to bf16 once."""
    E = _E()
    f = E.fused
    dtype = torch.bfloat16
    a = _rand((M, K), 96).to(dtype).cuda()
    w = (_rand((N, K), 97, (2.0 / K) ** 0.5) if not kn else _rand((K, N), 97, (2.0 / K) ** 0.5)).cuda()
    w16 = w.to(dtype)
    scale, shift = (_rand((K,), 98).abs() + 0.5).cuda(), _rand((K,), 99).cuda()
    outs = {}
    # the mirror path runs first, and a different GEMM in between overwrites the shared memory both paths
    # stage their weights into: a chunk one path forgets to write cannot be inherited from the other
    other_a, other_w = _rand((256, 64), 100).to(dtype).cuda(), _rand((64, 64), 101).cuda()
    other_out = torch.empty((256, 64), dtype=dtype, device="cuda")
    for mirror in (w16.data_ptr(), 0):
        out = torch.full((M, N), float("nan"), dtype=dtype, device="cuda")
        _call("ehgr_pw_gemm_w16", ctypes.byref(f.op_affine(a, scale, shift, True)), w.data_ptr(), mirror, kn, out.data_ptr(), 0, 0,
              M, K, N, 1, 2, _sp())
        for _ in range(3):
            _call("ehgr_pw_gemm", ctypes.byref(f.op_plain(other_a)), other_w.data_ptr(), 0, other_out.data_ptr(), 0, 0,
                  256, 64, 64, 1, 2, _sp())
        torch.cuda.synchronize()
        outs[mirror != 0] = out
    assert torch.isfinite(outs[True].float()).all()
    assert torch.equal(outs[False], outs[True])
    aa = torch.clamp(a.double() * scale.double() + shift.double(), 0, 6).to(dtype).double()
    want = aa @ (w16.double().t() if not kn else w16.double())
    assert rel_err(outs[True].cpu(), want.cpu()) < 1e-2


def test_pw_gemm_tcgen05_shift_prologue_full_size():
    """BASELINE config #2 size of the largest shifted layer (24->144 at 56x56, 256 frames): the shift
    fused into the GEMM A-load equals shift-then-GEMM."""
    E = _E()
    f = E.fused
    nt, hw, K, N, T = 256, 56 * 56, 24, 144, 8
    M = nt * hw
    gcuda = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((M, K), device="cuda", generator=gcuda).to(torch.bfloat16)
    w = torch.randn((N, K), device="cuda", generator=gcuda) * 0.3
    out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    _call("ehgr_pw_gemm", ctypes.byref(f.op_shift(x, T, K // 8, hw, 1)), w.data_ptr(), 0, out.data_ptr(), 0, 0, M, K, N, 1, 2, _sp())
    xs = torch.empty_like(x)
    _call("ehgr_row_apply", ctypes.byref(f.op_shift(x, T, K // 8, hw, 1)), 0, xs.data_ptr(), M, K, 1, _sp())
    want = xs.float() @ w.to(torch.bfloat16).float().t()
    assert rel_err(out.float().cpu(), want.cpu()) < 1e-2


# ---------------------------------------------------------------------------------------------
# depthwise 3x3
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nt,h,w,c,stride", [(3, 9, 7, 32, 1), (2, 8, 8, 96, 2), (2, 7, 7, 960, 1), (2, 14, 14, 144, 2), (1, 1, 1, 16, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dw_fwd_bwd(nt, h, w, c, stride, dtype):
    E = _E()
    f = E.fused
    x = _rand((nt, c, h, w), 20).to(dtype)
    wt = _rand((c, 1, 3, 3), 21, 0.4)
    scale, shift = _rand((c,), 22).abs() + 0.5, _rand((c,), 23)
    xr = _rows(x).contiguous().cuda()
    code = E._lib.dtype_code(xr)
    a64 = torch.clamp(x.double() * scale.double().view(1, c, 1, 1) + shift.double().view(1, c, 1, 1), 0, 6).requires_grad_(True)
    w64 = wt.double().requires_grad_(True)
    y64 = F.conv2d(a64, w64, stride=stride, padding=1, groups=c)
    ho, wo = y64.shape[2:]
    sc_d, sh_d, wt_d = scale.cuda(), shift.cuda(), wt.cuda()
    a_op = f.op_affine(xr, sc_d, sh_d, True)
    out = torch.empty((nt * ho * wo, c), dtype=dtype, device="cuda")
    stats = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    _call("ehgr_dw_fwd", ctypes.byref(a_op), wt_d.data_ptr(), out.data_ptr(), stats.data_ptr(), nt, h, w, c, stride, code, _sp())
    assert rel_err(out.cpu(), _rows(y64)) < TOL[dtype]
    stol = 1e-4 if dtype == torch.float32 else 1e-2   # bf16: staged activations and weights are rounded
    assert rel_err(stats[:c].cpu(), y64.sum((0, 2, 3))) < stol
    assert rel_err(stats[c:].cpu(), (y64 ** 2).sum((0, 2, 3))) < stol
    g = _rand(tuple(y64.shape), 24).to(dtype)
    y64.backward(g.double())
    gr = _rows(g).contiguous().cuda()
    da = torch.empty((nt * h * w, c), dtype=dtype, device="cuda")
    _call("ehgr_dw_dgrad", ctypes.byref(f.op_plain(gr)), wt_d.data_ptr(), da.data_ptr(), nt, h, w, c, stride, code, _sp())
    assert rel_err(da.cpu(), _rows(a64.grad)) < TOL[dtype]
    dw = torch.zeros((c, 9), dtype=torch.float32, device="cuda")
    _call("ehgr_dw_wgrad", ctypes.byref(f.op_plain(gr)), ctypes.byref(a_op), dw.data_ptr(), nt, h, w, c, stride, code, _sp())
    assert rel_err(dw.cpu(), w64.grad.view(c, 9)) < max(TOL[dtype], 2e-5)
    # fused backward (shared-memory tiled): same two results from one kernel
    da2 = torch.full((nt * h * w, c), float("nan"), dtype=dtype, device="cuda")
    dw2 = torch.zeros((c, 9), dtype=torch.float32, device="cuda")
    _call("ehgr_dw_bwd", ctypes.byref(f.op_plain(gr)), ctypes.byref(a_op), wt_d.data_ptr(), da2.data_ptr(), dw2.data_ptr(),
          nt, h, w, c, stride, code, _sp())
    assert rel_err(da2.cpu(), _rows(a64.grad)) < TOL[dtype]
    assert rel_err(dw2.cpu(), w64.grad.view(c, 9)) < max(TOL[dtype], 2e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dw_tiled_ragged_tiles(dtype):
    """Spatial sizes that do not divide the tile (partial tiles, odd sizes under stride 2)."""
    E = _E()
    f = E.fused
    for (nt, h, w, c, stride) in [(2, 17, 23, 48, 1), (2, 17, 23, 48, 2), (1, 30, 9, 96, 2), (3, 15, 15, 24, 1)]:
        x = _rand((nt, c, h, w), 25).to(dtype)
        wt = _rand((c, 1, 3, 3), 26, 0.4)
        a64 = x.double().requires_grad_(True)
        w64 = wt.double().requires_grad_(True)
        y64 = F.conv2d(a64, w64, stride=stride, padding=1, groups=c)
        ho, wo = y64.shape[2:]
        xr, wt_d = _rows(x).contiguous().cuda(), wt.cuda()
        code = E._lib.dtype_code(xr)
        out = torch.full((nt * ho * wo, c), float("nan"), dtype=dtype, device="cuda")
        stats = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
        _call("ehgr_dw_fwd", ctypes.byref(f.op_plain(xr)), wt_d.data_ptr(), out.data_ptr(), stats.data_ptr(), nt, h, w, c, stride, code, _sp())
        assert rel_err(out.cpu(), _rows(y64)) < TOL[dtype]
        assert rel_err(stats[:c].cpu(), y64.sum((0, 2, 3))) < (1e-3 if dtype == torch.float32 else 1e-2)
        g = _rand(tuple(y64.shape), 27).to(dtype)
        y64.backward(g.double())
        gr = _rows(g).contiguous().cuda()
        da = torch.full((nt * h * w, c), float("nan"), dtype=dtype, device="cuda")
        dw = torch.zeros((c, 9), dtype=torch.float32, device="cuda")
        _call("ehgr_dw_bwd", ctypes.byref(f.op_plain(gr)), ctypes.byref(f.op_plain(xr)), wt_d.data_ptr(), da.data_ptr(), dw.data_ptr(),
              nt, h, w, c, stride, code, _sp())
        assert rel_err(da.cpu(), _rows(a64.grad)) < TOL[dtype], (h, w, stride)
        assert rel_err(dw.cpu(), w64.grad.view(c, 9)) < max(TOL[dtype], 2e-5), (h, w, stride)


# ---------------------------------------------------------------------------------------------
# stem
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("hw,cout", [((32, 32), 32), ((17, 23), 32), ((50, 46), 32), ((224, 224), 32), ((20, 18), 16)])
def test_stem(dtype, hw, cout):
    """cout = 32 takes the banded shared-memory kernels, anything else the generic ones."""
    E = _E()
    f = E.fused
    h, w = hw
    nt = 3
    x = _rand((nt, 3, h, w), 30)
    wt = _rand((cout, 3, 3, 3), 31, 0.3)
    x64, w64 = x.double(), wt.double().requires_grad_(True)
    y64 = F.conv2d(x64, w64, stride=2, padding=1)
    ho, wo = y64.shape[2:]
    out = torch.empty((nt * ho * wo, cout), dtype=dtype, device="cuda")
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    xd, wt_d = x.cuda(), wt.cuda()
    code = E._lib.dtype_code(out)
    _call("ehgr_stem_fwd", xd.data_ptr(), wt_d.data_ptr(), out.data_ptr(), stats.data_ptr(), nt, h, w, cout, 0, code, _sp())
    assert rel_err(out.cpu(), _rows(y64)) < TOL[dtype]
    assert rel_err(stats[:cout].cpu(), y64.sum((0, 2, 3))) < 1e-4
    assert rel_err(stats[cout:].cpu(), (y64 * y64).sum((0, 2, 3))) < 1e-4
    g = _rand(tuple(y64.shape), 32).to(dtype)
    y64.backward(g.double())
    dw = torch.zeros((cout, 27), dtype=torch.float32, device="cuda")
    g_d = _rows(g).contiguous().cuda()
    _call("ehgr_stem_wgrad", ctypes.byref(f.op_plain(g_d)), xd.data_ptr(), dw.data_ptr(), nt, h, w, cout, 0,
          code, _sp())
    assert rel_err(dw.cpu(), w64.grad.view(cout, 27)) < max(TOL[dtype], 2e-5)


# ---------------------------------------------------------------------------------------------
# BatchNorm bookkeeping against nn.BatchNorm2d semantics
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("training", [True, False])
def test_bn_forward_backward_against_torch(training):
    E = _E()
    f = E.fused
    M, C = 5 * 6 * 7, 40
    x = _rand((5, C, 6, 7), 40) * 2 + 0.5
    bn = torch.nn.BatchNorm2d(C).double()
    bn.weight.data = _rand((C,), 41).double().abs() + 0.5
    bn.bias.data = _rand((C,), 42).double()
    bn.running_mean.data = _rand((C,), 43).double() * 0.2
    bn.running_var.data = _rand((C,), 44).double().abs() + 0.5
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    bn.train(training)
    x64 = x.double().requires_grad_(True)
    y = F.relu6(bn(x64))
    g = _rand(tuple(y.shape), 45)
    y.backward(g.double())

    xr = _rows(x).contiguous().cuda()
    stats = torch.stack([xr.double().sum(0), (xr.double() ** 2).sum(0)]).reshape(-1).contiguous()
    gamma, beta = bn.weight.data.float().cuda(), bn.bias.data.float().cuda()
    rm, rv = rm0.float().cuda(), rv0.float().cuda()
    vec = torch.empty((4, C), dtype=torch.float32, device="cuda")
    _call("ehgr_bn_finalize", stats.data_ptr(), M, gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5,
          int(training), vec[0].data_ptr(), vec[1].data_ptr(), vec[2].data_ptr(), vec[3].data_ptr(), C, _sp())
    out = torch.empty_like(xr)
    _call("ehgr_row_apply", ctypes.byref(f.op_affine(xr, vec[0], vec[1], True)), 0, out.data_ptr(), M, C, 0, _sp())
    assert rel_err(out.cpu(), _rows(y)) < 1e-5
    assert rel_err(rm.cpu(), bn.running_mean) < 1e-6 and rel_err(rv.cpu(), bn.running_var) < 1e-6
    # backward
    gr = _rows(g).contiguous().cuda()
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    _call("ehgr_bn_bwd_reduce", gr.data_ptr(), xr.data_ptr(), vec[0].data_ptr(), vec[1].data_ptr(), 1, sums.data_ptr(), M, C, 0, _sp())
    coef = torch.empty((3, C), dtype=torch.float32, device="cuda")
    dgb = torch.empty((2, C), dtype=torch.float32, device="cuda")
    _call("ehgr_bn_bwd_finalize", sums.data_ptr(), M, gamma.data_ptr(), vec[2].data_ptr(), vec[3].data_ptr(), int(training),
          coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), C, _sp())
    dx = torch.empty_like(xr)
    _call("ehgr_row_apply", ctypes.byref(f.op_bnbwd(gr, xr, coef[0], coef[1], coef[2], vec[0], vec[1], True)), 0, dx.data_ptr(),
          M, C, 0, _sp())
    assert rel_err(dx.cpu(), _rows(x64.grad)) < 1e-5
    assert rel_err(dgb[0].cpu(), bn.weight.grad) < 1e-5 and rel_err(dgb[1].cpu(), bn.bias.grad) < 1e-5


# ---------------------------------------------------------------------------------------------
# classifier head
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pool_and_fc_consensus(dtype):
    E = _E()
    nt, c, h, w, T, K = 16, 1280, 7, 7, 8, 83
    x = _rand((nt, c, h, w), 50).to(dtype)
    xd = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    lin = torch.nn.Linear(c, K).cuda()
    pooled = E.fused.global_avg_pool(xd)
    logits = E.fused.fc_consensus(pooled, lin, T)
    x64 = x.double().requires_grad_(True)
    w64, b64 = lin.weight.detach().cpu().double().requires_grad_(True), lin.bias.detach().cpu().double().requires_grad_(True)
    p64 = x64.mean(3).mean(2)
    l64 = F.linear(p64, w64, b64).view(-1, T, K).mean(1)
    assert rel_err(pooled.detach().cpu(), p64) < TOL[dtype]
    assert rel_err(logits.detach().cpu(), l64) < TOL[dtype]
    g = _rand(tuple(l64.shape), 51)
    logits.backward(g.cuda())
    l64.backward(g.double())
    assert rel_err(xd.grad.cpu(), x64.grad) < TOL[dtype]
    assert rel_err(lin.weight.grad.cpu(), w64.grad) < max(TOL[dtype], 1e-5)
    assert rel_err(lin.bias.grad.cpu(), b64.grad) < 1e-5


# ---------------------------------------------------------------------------------------------
# loss heads: oracle + reference golden fixtures
# ---------------------------------------------------------------------------------------------
def test_losses_match_reference_fixtures():
    from conftest import GOLDEN
    E = _E()
    z = np.load(GOLDEN / "losses.npz")
    labels = torch.from_numpy(z["sd_labels"]).cuda()
    logits = [torch.from_numpy(z[f"sd_logits{i}"]).cuda().requires_grad_(True) for i in range(4)]
    feats = [torch.from_numpy(z[f"sd_feat{i}"]).cuda().requires_grad_(True) for i in range(4)]
    total, terms = E.losses.sd_loss(logits, feats, labels, 0.1, 1e-6, 3.0)
    assert abs(total.item() - float(z["sd_total"])) < 2e-5 * abs(float(z["sd_total"]))
    assert np.allclose(terms.cpu().numpy(), z["sd_terms"], rtol=2e-5)
    total.backward()
    for i in range(4):
        assert rel_err(logits[i].grad.cpu(), torch.from_numpy(z[f"sd_glogits{i}"])) < 1e-5
        if i > 0:
            assert rel_err(feats[i].grad.cpu(), torch.from_numpy(z[f"sd_gfeat{i}"])) < 1e-5
    assert feats[0].grad is None

    lg = torch.from_numpy(z["mt_logits"]).cuda().requires_grad_(True)
    pred = torch.from_numpy(z["mt_pred"]).cuda().requires_grad_(True)
    depth = torch.from_numpy(z["mt_depth"].astype(np.float32)).cuda()
    loss, dl = E.losses.mtmm_loss(lg, labels, pred, depth)
    assert abs(loss.item() - float(z["mt_loss"])) < 2e-6 and abs(dl.item() - float(z["mt_depth_loss"])) < 1e-6
    (loss * 2.0).backward()     # exercises the incoming-gradient scaling
    assert rel_err(lg.grad.cpu() / 2, torch.from_numpy(z["mt_glogits"])) < 1e-5
    assert rel_err(pred.grad.cpu() / 2, torch.from_numpy(z["mt_gpred"])) < 1e-5


# ---------------------------------------------------------------------------------------------
# N4: uint8 frames normalised on the device == the reference's CPU ToTorchFormatTensor + GroupNormalize
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 4, 3, 32, 32), (1, 2, 3, 17, 20), (3, 1, 1, 56, 56)])
def test_normalize_u8_bit_exact(shape):
    E = _E()
    g = torch.Generator().manual_seed(5)
    x = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g)
    c = shape[2]
    mean, std = ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]) if c == 3 else (None, None)
    want = x.float().div(255)                                  # ToTorchFormatTensor(div=True)
    if mean is not None:                                        # GroupNormalize: t.sub_(m).div_(s) per channel
        for ci in range(c):
            want[:, :, ci].sub_(mean[ci]).div_(std[ci])
    got = E.train_step.normalize_u8(x.cuda(), mean, std)
    assert torch.equal(got.cpu(), want)
    got16 = E.train_step.normalize_u8(x.cuda(), mean, std, out_dtype=torch.bfloat16)
    assert torch.equal(got16.cpu(), want.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------
# round 2: TemporalPool kernel, EMA kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tp_a", "tp_b", "tp_c"])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_temporal_pool_matches_reference_fixture_bit_exact(name, layout):
    """TemporalPool.temporal_pool (models/temporal_shift.py:89-98) forward and backward against the fixture produced by
    the live reference (tests/golden/ema_pool.npz); the upstream gradient of the fixture is 0.5 + y."""
    E = _E()
    z = np.load(GOLDEN / "ema_pool.npz")
    nt, c, h, T = (int(v) for v in z[name + "_meta"])
    x = torch.from_numpy(z[name + "_x"]).cuda()
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y = E.TemporalPool.temporal_pool(x, T)
    want = torch.from_numpy(z[name + "_y"])
    assert tuple(y.shape) == tuple(want.shape) and torch.equal(y.cpu(), want)
    y.backward(0.5 + y.detach())
    assert torch.equal(x.grad.cpu(), torch.from_numpy(z[name + "_gx"]))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_temporal_pool_vector_path_against_torch(dtype):
    E = _E()
    g0 = torch.Generator().manual_seed(5)
    x = torch.randn(16, 32, 7, 7, generator=g0).to(dtype).cuda().requires_grad_(True)
    y = E.TemporalPool.temporal_pool(x, 8)
    xr = x.detach().float().clone().requires_grad_(True)
    yr = torch.nn.functional.max_pool3d(xr.view(2, 8, 32, 7, 7).transpose(1, 2), (3, 1, 1), (2, 1, 1), (1, 0, 0)).transpose(1, 2).reshape(8, 32, 7, 7)
    assert torch.equal(y.float(), yr)
    gup = torch.randn(8, 32, 7, 7, generator=g0).to(dtype).cuda()
    y.backward(gup)
    yr.backward(gup.float())
    assert rel_err(x.grad.float(), xr.grad) < (1e-6 if dtype == torch.float32 else 1e-2)


def test_ema_kernels_replay_the_reference_fixture_bit_exact():
    """ehgr_ema_update (float buffers and int64 counters) on the states recorded from the reference EMAWrapper
    (train_mtmm.py:110-128; tests/golden/ema_pool.npz): three updates at decay 0.9, every entry bit-identical."""
    z = np.load(GOLDEN / "ema_pool.npz")
    keys = [k[len("ema_init_"):] for k in z.files if k.startswith("ema_init_")]
    for k in keys:
        ema = torch.from_numpy(z["ema_init_" + k].copy()).cuda().reshape(-1)
        is_int = ema.dtype == torch.int64
        for step in range(3):
            m = torch.from_numpy(z[f"ema_model{step}_{k}"]).cuda().reshape(-1)
            _call("ehgr_ema_update", ema.data_ptr(), m.data_ptr(), ema.numel(), 0.9, int(is_int), _sp())
        assert torch.equal(ema.cpu(), torch.from_numpy(z["ema_final_" + k]).reshape(-1)), k
