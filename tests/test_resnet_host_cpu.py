"""N3 host logic on CPU: the orchestration of resnet_ops._ResNetFunction (which C-ABI entry point is called with which
operands, shapes and saved tensors; where every gradient lands) against PyTorch autograd over the same torchvision
modules.  The kernels are replaced by the numpy restatement of their header contracts (tests/abi_emulator.py) — the CUDA
kernels themselves are checked on the GPU (tests/test_resnet_gpu.py).  The reference path being mirrored:
torchvision.models.resnet50 built by models/models.py:108-117 with TemporalShift on every Bottleneck.conv1
(models/temporal_shift.py:101-146)."""
import contextlib
import io

import pytest
import torch
import torch.nn as nn

import abi_emulator
from conftest import check_grads_up_to_relu_flips, rel_err
from oracle import ref_oracle as O


@pytest.fixture()
def emulated(monkeypatch):
    import ehgr_b200
    monkeypatch.setattr(ehgr_b200._lib, "call", abi_emulator.call)
    monkeypatch.setattr(ehgr_b200._lib, "require_cuda", lambda *t: None)
    monkeypatch.setattr(ehgr_b200._lib, "stream_ptr", lambda device=None: 0)
    monkeypatch.setattr(ehgr_b200._lib, "on_gpu", lambda t: True)      # the wrappers take the library's path for host tensors
    # the oracle's differentiable shift stands in for the shift kernel in the eager (reference-side) run
    monkeypatch.setattr(ehgr_b200.TemporalShift, "shift",
                        staticmethod(lambda x, n_segment, fold_div=3, inplace=False: O.temporal_shift(x, n_segment, fold_div)))
    return ehgr_b200


def _net(E, layers, shift, n_segment, seed=0):
    from torchvision.models.resnet import Bottleneck, ResNet
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = ResNet(Bottleneck, layers)
        if shift:
            E.make_temporal_shift(net, n_segment, n_div=8, place='blockres')
    g = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():                               # non-trivial BatchNorm parameters and running statistics
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data = torch.rand(m.num_features, generator=g) + 0.5
            m.bias.data = torch.randn(m.num_features, generator=g) * 0.2
            m.running_mean.data = torch.randn(m.num_features, generator=g) * 0.1
            m.running_var.data = torch.rand(m.num_features, generator=g) + 0.5
    return net


def _eager(net, x):
    y = net.maxpool(net.relu(net.bn1(net.conv1(x))))
    t1 = net.layer1(y)
    t2 = net.layer2(t1)
    t3 = net.layer3(t2)
    return t1, t2, t3, net.layer4(t3)


def _set_mode(net, train_bn):
    """True / False: every BatchNorm in train / eval mode; "partial": TSN.train() with partial_bn (models/models.py:214-230) —
    every BatchNorm but the first frozen (eval mode, no gradient for its affine parameters)."""
    net.train(train_bn is not False)
    if train_bn == "partial":
        bns = [m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]
        for m in bns[1:]:
            m.eval()
            m.weight.requires_grad = False
            m.bias.requires_grad = False


@pytest.mark.parametrize("shift,train_bn,stem_gemm", [(True, True, False), (False, True, True), (True, False, True),
                                                      (True, "partial", False)])
def test_resnet_function_matches_autograd(emulated, monkeypatch, shift, train_bn, stem_gemm):
    E = emulated
    R = E.resnet_ops
    monkeypatch.setattr(R, "STEM_GEMM", stem_gemm)        # the 7x7 stem as a patch-matrix GEMM (the bf16 path) or the CUDA-core kernel
    T, size = 2, 48
    net = _net(E, [2, 1, 1, 1], shift, T)
    _set_mode(net, train_bn)
    ok, why = R.supported(net)
    assert ok, why
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2 * T, 3, size, size, generator=g)
    gouts = None

    # ---- reference: PyTorch autograd over the same modules, float64
    ref = _net(E, [2, 1, 1, 1], shift, T).double()
    _set_mode(ref, train_bn)
    outs_ref = _eager(ref, x.double())
    gouts = [torch.randn(o.shape, generator=g, dtype=torch.float64) / o.numel() ** 0.5 for o in outs_ref]
    torch.autograd.backward(outs_ref, gouts)

    # ---- ours: one autograd Function over the emulated C ABI, float32
    with E.fused.compute_dtype(torch.float32):
        outs = R.resnet_features(net, x, taps=(1, 2, 3))
    assert len(outs) == 4
    for o, r in zip(outs, outs_ref):
        assert o.shape == r.shape
        assert rel_err(o, r) < 2e-4
    torch.autograd.backward(outs, [q.float() for q in gouts])
    frozen = [k for k, p in net.named_parameters() if not p.requires_grad]
    assert (len(frozen) == 2 * (len([m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]) - 1)) == (train_bn == "partial")
    assert all(dict(net.named_parameters())[k].grad is None for k in frozen)
    check_grads_up_to_relu_flips(((k, p) for k, p in net.named_parameters() if not k.startswith("fc.") and p.requires_grad),
                                 {k: p.grad for k, p in ref.named_parameters() if not k.startswith("fc.") and p.requires_grad})
    # BatchNorm buffers follow nn.BatchNorm2d
    for (name, b), (_, br) in zip(net.named_buffers(), ref.named_buffers()):
        if b.dtype.is_floating_point:
            assert rel_err(b, br) < 1e-4, name
        else:
            assert int(b) == int(br), name


def test_single_output_and_tap_selection(emulated):
    E = emulated
    R = E.resnet_ops
    net = _net(E, [1, 1, 1, 1], True, 2)
    x = torch.randn(2, 3, 32, 32)
    with E.fused.compute_dtype(torch.float32):
        f = R.resnet_features(net, x)
        assert isinstance(f, torch.Tensor) and f.shape == (2, 2048, 1, 1)
        t2, f2 = R.resnet_features(net, x, taps=(2,))
    assert t2.shape == (2, 512, 4, 4) and f2.shape == (2, 2048, 1, 1)
    # only the final map gets a gradient: the tapped stage still back-propagates through the trunk
    f2.float().sum().backward()
    assert net.conv1.weight.grad is not None and net.conv1.weight.grad.abs().sum() > 0
    with pytest.raises(ValueError):
        R.resnet_features(net, x, taps=(4,))


def test_unsupported_trees_are_named(emulated):
    import torchvision
    E = emulated
    R = E.resnet_ops
    ok, why = R.supported(torchvision.models.resnet18())
    assert not ok and "BasicBlock" in why
    ok, why = R.supported(torchvision.models.resnext50_32x4d())
    assert not ok and "grouped" in why
    with contextlib.redirect_stdout(io.StringIO()):
        net = torchvision.models.resnet50()
        E.action.make_temporal_shift(net, 4, n_div=8)
    ok, why = R.supported(net)
    assert not ok and "Action" in why


def _wrapper(E, cls_mod, T, num_class, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return cls_mod.TSN(num_class, T, 'RGB', base_model='resnet50', pretrain=None, dropout=0.5, partial_bn=False, is_shift=True,
                           shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224, temporal_module='tsm',
                           print_spec=False, **kw)


def _no_dropout(model):
    model.train()
    for d in model.modules():
        if isinstance(d, nn.Dropout):
            d.eval()
    return model


def test_mtmm_wrapper_on_resnet50_matches_the_oracle(emulated):
    """tsn_mtmm.TSN.forward on TSM-ResNet-50 (backbone through resnet_ops, classifier head, depth decoder over 2048 channels)
    against oracle.resnet_mtmm_forward — the restatement pinned to the UNMODIFIED reference wrapper models_MTMM.TSN
    (tests/golden/resnet_wrappers.npz): outputs and every parameter gradient."""
    import torch.nn.functional as F
    E = emulated
    T, cls, size = 2, 6, 64
    sd0 = O.build_resnet_mtmm_state(cls, "tsm", seed=8)
    model = _wrapper(E, E.tsn_mtmm, T, cls, modal='rgb_depth')
    model.load_state_dict(sd0, strict=True)
    _no_dropout(model)
    rgb, depth, labels = O.synthetic_clip_batch(2, T, size, cls, seed=9)
    with E.fused.compute_dtype(torch.float32):
        logits, dpred = model(rgb)
    gt = F.interpolate(depth.view(-1, 1, size, size), tuple(dpred.shape[-2:]), mode='bilinear')
    (F.cross_entropy(logits, labels) + 0.01 * F.mse_loss(dpred, gt)).backward()
    sd64 = O.clone_state(sd0, dtype=torch.float64)
    ol, od = O.resnet_mtmm_forward(rgb.double(), sd64, T, "tsm", 8, True)
    (F.cross_entropy(ol, labels) + 0.01 * F.mse_loss(od, gt.double())).backward()
    assert tuple(dpred.shape) == tuple(od.shape) == (2 * T, 1, 16, 16)
    assert rel_err(logits, ol) < 1e-4 and rel_err(dpred, od) < 1e-4
    check_grads_up_to_relu_flips(model.named_parameters(), {k: v.grad for k, v in sd64.items() if v.is_floating_point()})


def test_sd_wrapper_on_resnet50_matches_the_oracle(emulated):
    """tsn_sd.TSN.forward on TSM-ResNet-50: layer1-3 taps of ONE resnet_ops pass feed the SepConv exit heads (fused chains
    of depthwise / pointwise stages), eight outputs as models/models_SD.py:431 — against oracle.resnet_sd_forward (pinned to
    the unmodified models_SD.TSN) with the SD loss of train_sd.py:227-265."""
    E = emulated
    T, cls, size = 2, 6, 64
    sd0 = O.build_resnet_sd_state(cls, "tsm", seed=8)
    model = _wrapper(E, E.tsn_sd, T, cls)
    model.load_state_dict(sd0, strict=True)
    _no_dropout(model)
    rgb, _, labels = O.synthetic_clip_batch(2, T, size, cls, seed=9)
    with E.fused.compute_dtype(torch.float32):
        outs = model(rgb)
    total, _ = O.sd_loss(outs[:4], outs[4:], labels, 0.1, 1e-6, 3.0)      # the loss as torch ops on OUR outputs
    total.backward()
    sd64 = O.clone_state(sd0, dtype=torch.float64)
    oouts = O.resnet_sd_forward(rgb.double(), sd64, T, "tsm", 8, True)
    ototal, _ = O.sd_loss(oouts[:4], oouts[4:], labels, 0.1, 1e-6, 3.0)
    ototal.backward()
    assert [tuple(o.shape) for o in outs] == [tuple(o.shape) for o in oouts] == [(2, cls)] * 4 + [(2 * T, 2048, 1, 1)] * 4
    for i, (a, b) in enumerate(zip(outs, oouts)):
        assert rel_err(a, b) < 2e-4, i
    assert abs(total.item() - ototal.item()) < 1e-4 * abs(ototal.item())
    check_grads_up_to_relu_flips(model.named_parameters(), {k: v.grad for k, v in sd64.items() if v.is_floating_point()})


def test_mtmm_sd_wrapper_on_resnet50_matches_the_oracle(emulated):
    """tsn_mtmm_sd.TSN.forward on TSM-ResNet-50 (the only backbone the reference's models_MTMM_SD.py supports): ONE
    resnet_ops pass with the max-pool output and layer1-3 tapped, exit heads, ConvTranspose decoders (library modules);
    ten outputs and the gradients of the combined loss (train_mtmm_sd.py:240-293) against the oracle."""
    import torch.nn.functional as F
    E = emulated
    T, cls, size = 2, 6, 64
    sd0 = O.build_resnet_mtmm_sd_state(cls, "tsm", seed=8)
    model = _wrapper(E, E.tsn_mtmm_sd, T, cls, modal='rgb_depth')
    model.load_state_dict(sd0, strict=True)
    _no_dropout(model)
    rgb, depth, labels = O.synthetic_clip_batch(2, T, size, cls, seed=9)
    with E.fused.compute_dtype(torch.float32):
        outs = model(rgb)
    assert len(outs) == 10 and tuple(outs[8].shape) == (2 * T, 1, size, size) and tuple(outs[9].shape) == (2 * T, 1, size // 4, size // 4)
    gt = F.interpolate(depth.view(-1, 1, size, size), (size // 4, size // 4), mode='bilinear')

    def total_of(o, gt_):       # combined loss + a term on the local decoder so that the tapped max-pool output gets a gradient
        t, _ = O.sd_loss(o[:4], o[4:8], labels, 0.1, 1e-6, 3.0)
        return t + 0.9 * 0.01 * F.mse_loss(o[9], gt_) + 0.01 * (o[8] ** 2).mean()
    total_of(outs, gt).backward()
    sd64 = O.clone_state(sd0, dtype=torch.float64)
    oouts = O.resnet_mtmm_sd_forward(rgb.double(), sd64, T, "tsm", 8, True)
    total_of(oouts, gt.double()).backward()
    for i, (a, b) in enumerate(zip(outs, oouts)):
        assert a.shape == b.shape and rel_err(a, b) < 2e-4, i
    check_grads_up_to_relu_flips(model.named_parameters(), {k: v.grad for k, v in sd64.items() if v.is_floating_point()})
