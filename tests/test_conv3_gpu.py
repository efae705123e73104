"""N2 — the MTMM depth decoder (models/models_MTMM.py:129-155) on the library's own kernels: the dense 3x3 convolution as
an implicit GEMM through the CONV3 row operand (forward, dgrad with the flipped filter, wgrad), the nearest-x2 upsample
folded into the gather and its adjoint, the depth head — each against PyTorch in fp64, and the whole decoder against the
fixture taken from the live reference's ``global_decoder`` (tests/golden/heads.npz, dec_a / dec_b)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_err
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def _E():
    import ehgr_b200
    return ehgr_b200


def _sp():
    return _E()._lib.stream_ptr(torch.device("cuda"))


def _nhwc(x, dtype):
    """[NT,C,H,W] cpu -> NHWC-contiguous device tensor [NT,H,W,C]."""
    return x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()


CASES = [  # nt, hs, ws (stored grid), cin, cout, up, affine
    (2, 7, 7, 64, 32, 0, False), (3, 3, 5, 32, 16, 1, True), (2, 14, 14, 64, 64, 1, True), (1, 28, 28, 32, 32, 0, True),
    (2, 7, 7, 256, 40, 0, False), (5, 4, 4, 8, 8, 1, False),
    # plain operands with 64-channel stages: the tcgen05 engine gathers them with TMA boxes (whole frames per tile when the
    # map is small, else cv_hbox image rows): ragged frame counts, 2 / 7 / 5-row boxes, several column chunks
    (3, 14, 14, 64, 32, 0, False), (2, 28, 28, 64, 16, 0, False), (5, 5, 3, 64, 16, 0, False), (1, 20, 24, 128, 24, 0, False),
    (3, 7, 7, 128, 320, 0, False), (2, 56, 56, 64, 8, 0, False),
    # 32-channel pixels: two taps per stage as SWIZZLE_64B boxes, half-filled last stage (K = 288)
    (2, 56, 56, 32, 32, 0, False), (3, 28, 28, 32, 64, 0, False), (5, 7, 7, 32, 16, 0, False)]


@pytest.mark.parametrize("nt,hs,ws,cin,cout,up,affine", CASES)
@pytest.mark.parametrize("dtype,engine", [(torch.float32, 1), (torch.bfloat16, 2), (torch.bfloat16, 1)])
def test_conv3_forward_dgrad_wgrad(nt, hs, ws, cin, cout, up, affine, dtype, engine):
    E = _E()
    f = E.fused
    g = torch.Generator().manual_seed(7 + cin + hs)
    x = torch.randn(nt, cin, hs, ws, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    scale, shift = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.3
    ho, wo = hs << up, ws << up
    M = nt * ho * wo
    code = 0 if dtype == torch.float32 else 1
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    xd = _nhwc(x, dtype)
    # ---- reference in fp64 (operands rounded as the kernels see them)
    xr = xd.cpu().double().permute(0, 3, 1, 2)
    a = torch.relu(xr * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)) if affine else xr
    if dtype == torch.bfloat16:
        a = a.to(dtype).double()
    if up:
        a = F.interpolate(a, scale_factor=2, mode="nearest")
    a.requires_grad_(True)
    wr = (w.to(dtype).double() if (dtype == torch.bfloat16 and engine == 2) else w.double()).requires_grad_(True)
    y = F.conv2d(a, wr, padding=1)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    gyd = _nhwc(gy.float(), dtype)
    y.backward(gyd.cpu().double().permute(0, 3, 1, 2))
    # ---- ours
    pdt = torch.float32 if engine == 1 else torch.bfloat16
    wf = torch.empty(cout * 9 * cin, dtype=pdt, device="cuda")
    wdd = torch.empty_like(wf)
    wdev = w.cuda()
    E._lib.call("ehgr_conv3_pack", wdev.data_ptr(), wf.data_ptr(), wdd.data_ptr(), cout, cin, 0 if engine == 1 else 1, _sp())
    sc, sh = scale.cuda(), shift.cuda()
    op = f.op_conv3(xd, sc if affine else None, sh if affine else None, 2, ho, wo, cin, up)
    out = torch.full((nt, ho, wo, cout), float("nan"), dtype=dtype, device="cuda")
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    w32, w16 = (wf, None) if engine == 1 else (wdev, wf)
    E._lib.call("ehgr_pw_gemm_w16", ctypes.byref(op), w32.data_ptr(), E._lib.ptr(w16), 0, out.data_ptr(), 0, stats.data_ptr(), M,
                9 * cin, cout, code, engine, _sp())
    got = out.cpu().double().permute(0, 3, 1, 2)
    assert rel_err(got, y.detach()) < tol
    assert rel_err(stats[:cout].cpu(), got.sum((0, 2, 3))) < (1e-4 if engine == 2 or dtype == torch.float32 else 1e-2)  # SIMT: statistics of the unrounded accumulators
    # dgrad: gradient w.r.t. the (upsampled) operand, then folded back through the upsample
    gup = torch.full((nt, ho, wo, cin), float("nan"), dtype=dtype, device="cuda")
    d32, d16 = (wdd, None) if engine == 1 else (wdev, wdd)
    E._lib.call("ehgr_pw_gemm_w16", ctypes.byref(f.op_conv3(gyd, None, None, 0, ho, wo, cout, 0)), d32.data_ptr(), E._lib.ptr(d16),
                0, gup.data_ptr(), 0, 0, M, 9 * cout, cin, code, engine, _sp())
    assert rel_err(gup.cpu().double().permute(0, 3, 1, 2), a.grad) < tol
    if up:
        gl = torch.empty((nt, hs, ws, cin), dtype=dtype, device="cuda")
        E._lib.call("ehgr_upsample2_bwd", gup.data_ptr(), gl.data_ptr(), nt, hs, ws, cin, code, _sp())
        want = F.avg_pool2d(gup.cpu().double().permute(0, 3, 1, 2), 2) * 4
        assert rel_err(gl.cpu().double().permute(0, 3, 1, 2), want) < (1e-6 if dtype == torch.float32 else 1e-2)
    # wgrad
    dwp = torch.zeros(cout * 9 * cin, dtype=torch.float32, device="cuda")
    E._lib.call("ehgr_pw_wgrad", ctypes.byref(f.op_plain(gyd)), ctypes.byref(op), dwp.data_ptr(), M, 9 * cin, cout, code, engine, _sp())
    dw = torch.zeros(cout, cin, 3, 3, dtype=torch.float32, device="cuda")
    E._lib.call("ehgr_conv3_unpack_grad", dwp.data_ptr(), dw.data_ptr(), cout, cin, _sp())
    assert rel_err(dw.cpu().double(), wr.grad) < tol


def test_conv3_full_size_layer_bf16():
    """The largest decoder layer of the benchmarked step (B=32: 256 frames, 1280 -> 256 @ 7x7, K = 11520): tcgen05 path
    against the fp32 SIMT engine on the same bf16 operands (multi-tile, streamed weights, 180 ring stages per tile)."""
    E = _E()
    f = E.fused
    g = torch.Generator().manual_seed(3)
    nt, hs, cin, cout = 64, 7, 1280, 256
    xd = (torch.randn(nt, hs, hs, cin, generator=g)).to(torch.bfloat16).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5).cuda()
    M = nt * hs * hs
    wf16 = torch.empty(cout * 9 * cin, dtype=torch.bfloat16, device="cuda")
    E._lib.call("ehgr_conv3_pack", w.data_ptr(), wf16.data_ptr(), 0, cout, cin, 1, _sp())
    wf32 = wf16.float()
    op = f.op_conv3(xd, None, None, 0, hs, hs, cin, 0)
    outs = []
    for engine, w32, w16 in ((2, w, wf16), (1, wf32, None)):
        out = torch.empty((nt, hs, hs, cout), dtype=torch.bfloat16, device="cuda")
        E._lib.call("ehgr_pw_gemm_w16", ctypes.byref(op), w32.data_ptr(), E._lib.ptr(w16), 0, out.data_ptr(), 0, 0, M, 9 * cin, cout,
                    1, engine, _sp())
        outs.append(out.float().cpu())
    assert rel_err(outs[0], outs[1].double()) < 1e-2
    ref = F.conv2d(xd.float().permute(0, 3, 1, 2), wf16.float().view(cout, 9, cin).permute(0, 2, 1).reshape(cout, cin, 3, 3), padding=1)
    assert rel_err(outs[0].permute(0, 3, 1, 2), ref.cpu().double()) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_depth_head(dtype):
    E = _E()
    f = E.fused
    g = torch.Generator().manual_seed(11)
    nt, c, h = 3, 32, 9
    x = torch.randn(nt, c, h, h, generator=g)
    xd = _nhwc(x, dtype).permute(0, 3, 1, 2).requires_grad_(True)          # logical NCHW, NHWC strides
    w = (torch.randn(1, c, 1, 1, generator=g) * 0.3).cuda().requires_grad_(True)
    b = torch.tensor([0.1]).cuda().requires_grad_(True)
    out = f._DepthHeadFunction.apply(xd, w, b)
    go = torch.randn(out.shape, generator=g).cuda()
    out.backward(go)
    xr = xd.detach().cpu().double().requires_grad_(True)
    wr, br = w.detach().cpu().double().requires_grad_(True), b.detach().cpu().double().requires_grad_(True)
    ref = torch.sigmoid(F.conv2d(xr, wr, br))
    ref.backward(go.cpu().double())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(out.cpu(), ref.detach()) < 1e-5
    assert rel_err(xd.grad.cpu().double(), xr.grad) < tol
    assert rel_err(w.grad.cpu(), wr.grad) < 1e-5 and rel_err(b.grad.cpu(), br.grad) < 1e-5


def _digest64(g):
    g = g.detach().double().flatten().cpu()
    idx = torch.linspace(0, g.numel() - 1, steps=min(64, g.numel())).long()
    return np.concatenate([[g.sum().item(), g.abs().sum().item()], g[idx].numpy()])


@pytest.mark.parametrize("name", ["dec_a", "dec_b"])
def test_depth_decoder_against_reference_fixture(name):
    """fused.depth_decoder on the reference's own global_decoder architecture (2048 input channels, the ResNet-50 width)
    against the live reference module's output, input gradient, parameter-gradient digests and running variance (fp32
    storage, exact-fp32 engine: 1e-5 / 3e-5)."""
    E = _E()
    z = np.load(GOLDEN / "heads.npz")
    h, n, train_bn = (int(v) for v in z[name + "_meta"])
    sd = {}
    O.decoder_state(sd, np.random.RandomState(601 + h), feat=2048, prefix="d")
    dec = E.tsn_mtmm.make_global_decoder(2048)
    dec.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
    dec = dec.cuda().train(bool(train_bn))
    x = torch.from_numpy(z[name + "_x"].astype(np.float32)).cuda().requires_grad_(True)
    l0 = E._lib.launch_count()
    with E.fused.compute_dtype(torch.float32):
        y = E.fused.depth_decoder(dec, x)
    assert tuple(y.shape) == (n, 1, 8 * h, 8 * h)
    y.backward(torch.from_numpy(z[name + "_g"]).cuda())
    assert E._lib.launch_count() - l0 >= 30                   # our kernels ran (no library convolution path)
    assert rel_err(y, torch.from_numpy(z[name + "_y"])) < 1e-5
    assert rel_err(x.grad, torch.from_numpy(z[name + "_gx"])) < 3e-5
    params = dict(dec.named_parameters())
    for k in z.files:
        if k.startswith(name + "_gdig_"):
            got, ref = _digest64(params[k[len(name + "_gdig_"):]].grad), z[k]
            assert np.abs(got[2:] - ref[2:]).max() <= 3e-5 * max(np.abs(ref[2:]).max(), 1e-6), k
    if train_bn:
        assert rel_err(dec[13].running_var, torch.from_numpy(z[name + "_rv13"])) < 1e-5


def test_depth_decoder_bf16_against_fp32_engine():
    """bf16 storage + tcgen05 against the exact-fp32 engine on the MobileNetV2 decoder (1280 channels): output 2e-2."""
    E = _E()
    torch.manual_seed(5)
    dec = E.tsn_mtmm.make_global_decoder(1280).cuda().train()
    x = torch.randn(8, 1280, 7, 7, device="cuda")
    outs, grads = [], []
    for dt in (torch.float32, torch.bfloat16):
        xi = x.clone().requires_grad_(True)
        for m in dec.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.reset_running_stats()
        dec.zero_grad()
        with E.fused.compute_dtype(dt):
            y = E.fused.depth_decoder(dec, xi)
            y.square().mean().backward()
        outs.append(y.detach().float().cpu())
        grads.append(dec[0].weight.grad.detach().cpu().clone())
    assert rel_err(outs[1], outs[0].double()) < 2e-2
    # the weight gradient of the first layer passes four batch-statistics BatchNorms in bf16: direction, not digits
    cos = F.cosine_similarity(grads[1].flatten().double(), grads[0].flatten().double(), dim=0).item()
    assert cos > 0.97, cos


def test_conv3_operand_is_refused_outside_the_gemm_family():
    E = _E()
    x = torch.zeros(2, 4, 4, 8, device="cuda")
    out = torch.empty_like(x)
    op = E.fused.op_conv3(x, None, None, 0, 4, 4, 8, 0)
    st = E._lib.lib().ehgr_row_apply(ctypes.byref(op), None, out.data_ptr(), 32, 8, 0, None)
    assert st == -5
