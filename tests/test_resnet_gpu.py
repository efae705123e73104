"""N3 on the GPU — the ResNet-50 bottleneck path on the library's kernels.

Kernel level (csrc/resnet.cu): 7x7/2 stem forward + statistics + weight gradient, max-pool forward / backward, even-pixel
subsample and its adjoint, bn+add+relu and the ReLU mask — each against PyTorch in fp64 on the operands the kernel sees
(1e-5 for fp32 storage, 2e-2 for bf16).  Network level: torchvision resnet50 + TemporalShift on every conv1 (what
models/models.py:108-117 + models/temporal_shift.py:101-146 build) through resnet_ops / TSN.forward against the oracle
(oracle/ref_oracle.py resnet_*; pinned to the live torchvision + reference pair by tests/golden/resnet.npz)."""
import contextlib
import ctypes
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_err
from oracle import ref_oracle as O
from test_oracle_golden import RESNET_FIXTURE, resnet_fixture_gouts

pytestmark = pytest.mark.gpu

DTYPES = [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)]


def _E():
    import ehgr_b200
    return ehgr_b200


def _sp():
    return _E()._lib.stream_ptr(torch.device("cuda"))


def _code(dtype):
    return 0 if dtype == torch.float32 else 1


def _nhwc(x, dtype):
    """[F,C,H,W] cpu -> NHWC-contiguous device tensor [F,H,W,C]."""
    return x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()


def _back(t):
    """device NHWC [F,H,W,C] -> cpu fp64 [F,C,H,W]."""
    return t.detach().cpu().double().permute(0, 3, 1, 2)


@pytest.mark.parametrize("frames,h,w", [(3, 38, 38), (2, 64, 48), (1, 7, 9)])
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_stem7_forward_statistics_wgrad(frames, h, w, dtype, tol):
    E = _E()
    g = torch.Generator().manual_seed(h)
    x = torch.randn(frames, 3, h, w, generator=g)
    wt = torch.randn(64, 3, 7, 7, generator=g) * (2.0 / 147) ** 0.5
    xd = x.to(dtype).cuda()
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.full((frames, ho, wo, 64), float("nan"), dtype=dtype, device="cuda")
    stats = torch.zeros(128, dtype=torch.float64, device="cuda")
    E._lib.call("ehgr_stem7_fwd", xd.data_ptr(), wt.cuda().data_ptr(), out.data_ptr(), stats.data_ptr(), frames, h, w, 64,
                _code(dtype), _code(dtype), _sp())
    xr = xd.cpu().double().requires_grad_(True)
    wr = wt.double().requires_grad_(True)
    y = F.conv2d(xr, wr, stride=2, padding=3)
    assert tuple(y.shape[2:]) == (ho, wo)
    got = _back(out)
    assert rel_err(got, y) < tol
    assert rel_err(stats[:64].cpu(), got.sum((0, 2, 3))) < 1e-5
    assert rel_err(stats[64:].cpu(), (got * got).sum((0, 2, 3))) < 1e-5
    # weight gradient from a plain d(raw) operand
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    gyd = _nhwc(gy.float(), dtype)
    y.backward(_back(gyd))
    dw = torch.zeros(64, 3, 7, 7, dtype=torch.float32, device="cuda")
    E._lib.call("ehgr_stem7_wgrad", ctypes.byref(E.fused.op_plain(gyd)), xd.data_ptr(), dw.data_ptr(), frames, h, w, 64, _code(dtype),
                _code(dtype), _sp())
    assert rel_err(dw.cpu(), wr.grad) < 1e-5          # fp32 accumulation of the operands as stored


@pytest.mark.parametrize("frames,h,w", [(3, 38, 38), (2, 64, 48), (1, 7, 9)])
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_stem7_as_gemm(frames, h, w, dtype, tol):
    """The bf16 path of the stem: patch matrix (ehgr_stem7_im2col) x packed filter on the pointwise GEMM kernels, forward with
    statistics and weight gradient, against F.conv2d in fp64; the patch matrix itself against F.unfold (bit-exact copy)."""
    E = _E()
    KP = E.resnet_ops.STEM_KP
    g = torch.Generator().manual_seed(h + 1)
    x = torch.randn(frames, 3, h, w, generator=g)
    wt = (torch.randn(64, 3, 7, 7, generator=g) * (2.0 / 147) ** 0.5).cuda()
    xd = x.to(dtype).cuda()
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    M = frames * ho * wo
    A = torch.full((M, KP), float("nan"), dtype=dtype, device="cuda")
    E._lib.call("ehgr_stem7_im2col", xd.data_ptr(), A.data_ptr(), frames, h, w, KP, _code(dtype), _code(dtype), _sp())
    want = F.unfold(xd.cpu().double(), 7, padding=3, stride=2).transpose(1, 2).reshape(M, 147)
    assert torch.equal(A[:, :147].cpu().double(), want) and bool((A[:, 147:] == 0).all())
    wp32 = torch.full((64, KP), float("nan"), dtype=torch.float32, device="cuda")
    wp16 = torch.full((64, KP), float("nan"), dtype=torch.bfloat16, device="cuda")
    E._lib.call("ehgr_stem7_pack", wt.data_ptr(), wp32.data_ptr(), 64, KP, 0, _sp())
    E._lib.call("ehgr_stem7_pack", wt.data_ptr(), wp16.data_ptr(), 64, KP, 1, _sp())
    assert torch.equal(wp32[:, :147], wt.view(64, 147)) and bool((wp32[:, 147:] == 0).all())
    assert torch.equal(wp16, wp32.to(torch.bfloat16))
    engine = 1 if dtype == torch.float32 else 2
    out = torch.full((M, 64), float("nan"), dtype=dtype, device="cuda")
    stats = torch.zeros(128, dtype=torch.float64, device="cuda")
    E._lib.call("ehgr_pw_gemm_bn", ctypes.byref(E.fused.op_plain(A)), wp32.data_ptr(), wp16.data_ptr() if engine == 2 else 0, 0,
                out.data_ptr(), 0, stats.data_ptr(), M, KP, 64, _code(dtype), engine, None, _sp())
    wr = (wt.cpu().to(dtype).double() if engine == 2 else wt.cpu().double()).requires_grad_(True)
    y = F.conv2d(xd.cpu().double(), wr, stride=2, padding=3)
    got = out.view(frames, ho, wo, 64).cpu().double().permute(0, 3, 1, 2)
    assert rel_err(got, y) < tol
    assert rel_err(stats[:64].cpu(), got.sum((0, 2, 3))) < 1e-4
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    gyd = _nhwc(gy.float(), dtype)
    y.backward(_back(gyd))
    dwp = torch.zeros(64 * KP, dtype=torch.float32, device="cuda")
    E._lib.call("ehgr_pw_wgrad", ctypes.byref(E.fused.op_plain(gyd)), ctypes.byref(E.fused.op_plain(A)), dwp.data_ptr(), M, KP, 64,
                _code(dtype), engine, _sp())
    dw = torch.zeros(64, 3, 7, 7, dtype=torch.float32, device="cuda")
    E._lib.call("ehgr_stem7_unpack_grad", dwp.data_ptr(), dw.data_ptr(), 64, KP, _sp())
    assert rel_err(dw.cpu(), wr.grad) < tol


@pytest.mark.parametrize("frames,h,w,c", [(3, 16, 16, 64), (2, 7, 9, 64), (2, 12, 6, 24), (1, 1, 1, 8)])
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_maxpool3_forward_backward(frames, h, w, c, dtype, tol):
    E = _E()
    g = torch.Generator().manual_seed(c + h)
    raw = torch.randn(frames, c, h, w, generator=g)
    scale, shift = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    rd = _nhwc(raw, dtype)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.full((frames, ho, wo, c), float("nan"), dtype=dtype, device="cuda")
    idx = torch.full((frames, ho, wo, c), 255, dtype=torch.uint8, device="cuda")
    sc, sh = scale.cuda(), shift.cuda()
    E._lib.call("ehgr_maxpool3_fwd", ctypes.byref(E.fused.op_affine(rd, sc, sh, 2)), y.data_ptr(), idx.data_ptr(), frames, h, w, c,
                _code(dtype), _sp())
    # the activation exactly as the kernel forms it: fp32 fma of the stored raw values, ReLU
    a = torch.relu(torch.addcmul(shift.view(1, -1, 1, 1), _back(rd).float(), scale.view(1, -1, 1, 1))).double().requires_grad_(True)
    yr = F.max_pool2d(a, 3, 2, 1)
    assert tuple(yr.shape[2:]) == (ho, wo)
    assert rel_err(_back(y), yr) < (1e-6 if dtype == torch.float32 else 4e-3)      # bf16: one rounding of the stored maximum
    assert int(idx.max()) <= 8
    gy = torch.randn(yr.shape, generator=g, dtype=torch.float64)
    gyd = _nhwc(gy.float(), dtype)
    yr.backward(_back(gyd))
    gx = torch.full((frames, h, w, c), float("nan"), dtype=dtype, device="cuda")
    E._lib.call("ehgr_maxpool3_bwd", gyd.data_ptr(), idx.data_ptr(), gx.data_ptr(), frames, h, w, c, _code(dtype), _sp())
    # ties only occur between zeros of the ReLU, whose gradient the ReLU mask removes: compare where the activation is positive
    pos = (a.detach() > 0).double()
    assert rel_err(_back(gx) * pos, a.grad * pos) < tol
    # conservation: every output gradient lands on exactly one input element
    assert abs(_back(gx).sum().item() - _back(gyd).sum().item()) < 1e-3 * _back(gyd).abs().sum().item()


@pytest.mark.parametrize("frames,h,w,c", [(3, 8, 8, 64), (2, 7, 5, 32), (1, 14, 14, 2048), (5, 1, 3, 8)])
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_subsample2_and_adjoint(frames, h, w, c, dtype, tol):
    E = _E()
    g = torch.Generator().manual_seed(c + w)
    x = torch.randn(frames, c, h, w, generator=g)
    xd = _nhwc(x, dtype)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.full((frames, ho, wo, c), float("nan"), dtype=dtype, device="cuda")
    stats = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    E._lib.call("ehgr_subsample2_fwd", xd.data_ptr(), y.data_ptr(), stats.data_ptr(), frames, h, w, c, _code(dtype), _sp())
    want = _back(xd)[:, :, ::2, ::2]
    assert torch.equal(_back(y), want)                       # a copy: bit-exact
    assert rel_err(stats[:c].cpu(), want.sum((0, 2, 3))) < 1e-5 and rel_err(stats[c:].cpu(), (want * want).sum((0, 2, 3))) < 1e-5
    y2 = torch.empty_like(y)
    E._lib.call("ehgr_subsample2_fwd", xd.data_ptr(), y2.data_ptr(), 0, frames, h, w, c, _code(dtype), _sp())      # no statistics
    assert torch.equal(y2, y)
    gx = torch.full((frames, h, w, c), float("nan"), dtype=dtype, device="cuda")
    E._lib.call("ehgr_subsample2_bwd", y.data_ptr(), gx.data_ptr(), frames, h, w, c, _code(dtype), _sp())
    full = torch.zeros_like(_back(xd))
    full[:, :, ::2, ::2] = want
    assert torch.equal(_back(gx), full)


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_bn_add_relu_and_mask(dtype, tol):
    E = _E()
    g = torch.Generator().manual_seed(2)
    m, c = 777, 256
    raw, add = torch.randn(m, c, generator=g).to(dtype).cuda(), torch.randn(m, c, generator=g).to(dtype).cuda()
    scale, shift = (torch.rand(c, generator=g) + 0.5).cuda(), (torch.randn(c, generator=g) * 0.3).cuda()
    out = torch.empty_like(raw)
    E._lib.call("ehgr_bn_add_relu", raw.data_ptr(), scale.data_ptr(), shift.data_ptr(), add.data_ptr(), out.data_ptr(), m, c,
                _code(dtype), _sp())
    want = torch.relu(raw.double() * scale.double() + shift.double() + add.double())
    assert rel_err(out, want) < (1e-6 if dtype == torch.float32 else 4e-3)
    gy = torch.randn(m, c, generator=g).to(dtype).cuda()
    gz = torch.empty_like(gy)
    E._lib.call("ehgr_relu_bwd", gy.data_ptr(), out.data_ptr(), gz.data_ptr(), m * c, _code(dtype), _sp())
    assert torch.equal(gz, torch.where(out > 0, gy, torch.zeros_like(gy)))


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _tsn(E, cls, temporal, T, num_class, **kw):
    with _quiet():
        return cls(num_class, T, 'RGB', base_model='resnet50', pretrain=None, dropout=0.5, partial_bn=False,
                   is_shift=(temporal == "tsm"), shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224,
                   temporal_module='tsm', print_spec=False, **kw)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_resnet50_tsm_stage_outputs_against_fixture_and_oracle(dtype, tol):
    """The four stage outputs of TSM-ResNet-50 (2 clips x 4 frames x 64^2, train-mode BatchNorm) and every parameter gradient:
    arbiter = the fp64 fixture of the live torchvision + reference pair / the fp64 oracle; yardstick = the oracle run on the
    GPU at the same precision (fp32 with TF32 off, bf16 autocast): ours <= max(3 x yardstick, tol)."""
    E = _E()
    cfg = RESNET_FIXTURE
    z = np.load(GOLDEN / "resnet.npz")
    sd0 = O.build_resnet_state(O.RESNET50_LAYERS, cfg["num_class"], "tsm", seed=cfg["seed"])
    model = _tsn(E, E.TSN, "tsm", cfg["T"], cfg["num_class"])
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    ok, why = E.resnet_ops.supported(model.base_model)
    assert ok, why
    rgb, _, _ = O.synthetic_clip_batch(cfg["clips"], cfg["T"], cfg["size"], cfg["num_class"], seed=cfg["in_seed"])
    x = rgb.view((-1, 3) + tuple(rgb.shape[-2:]))
    n0 = E._lib.launch_count()
    with E.fused.compute_dtype(dtype):
        taps = E.resnet_ops.resnet_features(model.base_model, x.cuda(), taps=(1, 2, 3))
    assert E._lib.launch_count() - n0 > 150           # 16 bottlenecks x (3-4 convolutions + BatchNorm bookkeeping), our kernels
    assert [tuple(t.shape[1:]) for t in taps] == [(256, 16, 16), (512, 8, 8), (1024, 4, 4), (2048, 2, 2)]
    gouts = resnet_fixture_gouts([t.shape for t in taps])
    torch.autograd.backward(taps, [q.to(t.dtype).cuda() for q, t in zip(gouts, taps)])
    # arbiter (fp64) and yardstick (same precision, PyTorch ops) on the GPU
    sd64 = {k: (v.detach().double().cuda().requires_grad_(v.requires_grad) if v.is_floating_point() else v.cuda())
            for k, v in O.clone_state(sd0).items()}
    t64 = O.resnet_features(x.double().cuda(), sd64, O.RESNET50_LAYERS, "tsm", cfg["T"], 8, True)
    torch.autograd.backward(t64, [q.cuda() for q in gouts])
    assert rel_err(t64[3], torch.from_numpy(z["tsm_train_layer4"])) < 1e-7           # the GPU fp64 arbiter is the fixture
    sdy = {k: (v.detach().cuda().requires_grad_(v.requires_grad) if v.is_floating_point() else v.cuda())
           for k, v in O.clone_state(sd0).items()}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
            ty = O.resnet_features(x.cuda(), sdy, O.RESNET50_LAYERS, "tsm", cfg["T"], 8, True)
        torch.autograd.backward(ty, [q.to(t.dtype).cuda() for q, t in zip(gouts, ty)])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    for i in range(4):
        ours, ref = rel_err(taps[i], t64[i]), rel_err(ty[i], t64[i])
        assert ours <= max(3.0 * ref, tol), (f"stage {i + 1}", ours, ref)
    gmax = max(v.grad.abs().max().item() for v in sd64.values() if v.is_floating_point() and v.grad is not None)
    named = {k: p for k, p in model.named_parameters() if not k.startswith("new_fc.")}
    ours = {k: (p.grad.double() - sd64[k].grad).abs().max().item() / gmax for k, p in named.items()}
    ref = {k: (sdy[k].grad.double() - sd64[k].grad).abs().max().item() / gmax for k in named}
    worst_ref = max(ref.values())
    for k in named:
        assert ours[k] <= max(3.0 * worst_ref, tol), (k, ours[k], ref[k], worst_ref)
    # running statistics follow nn.BatchNorm2d
    msd = model.state_dict()
    for k in ("base_model.bn1.running_mean", "base_model.layer2.0.downsample.1.running_var", "base_model.layer4.2.bn3.running_var"):
        assert rel_err(msd[k], sd64[k]) < (1e-4 if dtype == torch.float32 else 3e-2), k
    assert int(msd["base_model.layer3.5.bn2.num_batches_tracked"]) == 1


@pytest.mark.parametrize("temporal", ["tsm", "none"])
def test_tsn_resnet50_forward_backward_runs_on_own_kernels(temporal):
    """TSN.forward (models/models.py:323-356) with a ResNet-50 base: logits, loss and gradients against the fp64 oracle,
    fp32 storage; no torchvision convolution may run (the library's launch counter accounts for the step)."""
    E = _E()
    cfg = RESNET_FIXTURE
    sd0 = O.build_resnet_state(O.RESNET50_LAYERS, cfg["num_class"], temporal, seed=cfg["seed"])
    model = _tsn(E, E.TSN, temporal, cfg["T"], cfg["num_class"])
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    for d in model.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    rgb, _, labels = O.synthetic_clip_batch(cfg["clips"], cfg["T"], cfg["size"], cfg["num_class"], seed=cfg["in_seed"])
    n0 = E._lib.launch_count()
    with E.fused.compute_dtype(torch.float32):
        logits = model(rgb.cuda())
        loss = F.cross_entropy(logits.float(), labels.cuda())
    loss.backward()
    assert E._lib.launch_count() - n0 > 400
    sd64 = {k: (v.detach().double().cuda().requires_grad_(v.requires_grad) if v.is_floating_point() else v.cuda())
            for k, v in O.clone_state(sd0).items()}
    ol = O.resnet_tsn_forward(rgb.double().cuda(), sd64, cfg["T"], O.RESNET50_LAYERS, temporal, 8, True)
    oloss = F.cross_entropy(ol, labels.cuda())
    oloss.backward()
    if temporal == "none":
        z = np.load(GOLDEN / "resnet.npz")
        assert rel_err(ol, torch.from_numpy(z["tsn_none_logits"])) < 1e-7            # arbiter == the live reference wrapper
    sdy = {k: (v.detach().cuda().requires_grad_(v.requires_grad) if v.is_floating_point() else v.cuda())
           for k, v in O.clone_state(sd0).items()}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yl = O.resnet_tsn_forward(rgb.cuda(), sdy, cfg["T"], O.RESNET50_LAYERS, temporal, 8, True)
        F.cross_entropy(yl, labels.cuda()).backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert rel_err(logits, ol) <= max(3.0 * rel_err(yl, ol), 2e-5)
    assert abs(loss.item() - oloss.item()) <= max(3.0 * abs(F.cross_entropy(yl, labels.cuda()).item() - oloss.item()), 2e-5)
    gmax = max(v.grad.abs().max().item() for v in sd64.values() if v.is_floating_point() and v.grad is not None)
    worst_ref = max((sdy[k].grad.double() - sd64[k].grad).abs().max().item() / gmax for k, _ in model.named_parameters())
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        e = (p.grad.double() - sd64[k].grad).abs().max().item() / gmax
        assert e <= max(3.0 * worst_ref, 2e-5), (k, e, worst_ref)


def test_mtmm_wrapper_takes_the_resnet_path():
    """models_MTMM.TSN on ResNet-50 (the reference's own configuration of that wrapper, models/models_MTMM.py:112-157):
    backbone through resnet_ops, layer4 [NT,2048,h,w] into the depth decoder on the implicit-GEMM kernels; shapes as the
    reference returns them, finite gradients everywhere, bf16 tensor-core path."""
    E = _E()
    T = 4
    rgb, depth, labels = O.synthetic_clip_batch(2, T, 128, 10, seed=3)
    mt = _tsn(E, E.tsn_mtmm.TSN, "tsm", T, 10, modal='rgb_depth').cuda().train()
    n0 = E._lib.launch_count()
    with E.fused.compute_dtype(torch.bfloat16):
        logits, dpred = mt(rgb.cuda())
    assert E._lib.launch_count() - n0 > 170
    assert tuple(logits.shape) == (2, 10) and tuple(dpred.shape) == (2 * T, 1, 32, 32)
    (F.cross_entropy(logits.float(), labels.cuda()) + dpred.float().mean()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in mt.parameters())


def test_sd_wrapper_takes_the_resnet_path():
    """models_SD.TSN on ResNet-50 (models/models_SD.py:214-253, 364-431): layer1-3 taps of resnet_ops feed the SepConv exit
    heads; eight outputs with the reference's shapes, finite gradients everywhere."""
    E = _E()
    T = 4
    rgb, _, labels = O.synthetic_clip_batch(2, T, 128, 10, seed=3)
    sdm = _tsn(E, E.tsn_sd.TSN, "tsm", T, 10).cuda().train()
    with E.fused.compute_dtype(torch.bfloat16):
        outs = sdm(rgb.cuda())
    assert [tuple(o.shape) for o in outs] == [(2, 10)] * 4 + [(2 * T, 2048, 1, 1)] * 4
    sum(o.float().sum() for o in outs).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in sdm.parameters())
