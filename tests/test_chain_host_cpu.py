"""Host logic of the MobileNetV2 path on CPU: fused._ChainFunction (stem + 17 InvertedResidual blocks + final 1x1, lazy
BatchNorm operands, TSM shift operands, residual gradients, parameter-gradient routing), the classifier head and the MTMM
depth decoder — run through the numpy / torch restatement of the C-ABI contracts (tests/abi_emulator.py) and compared with
the oracle (pinned to the live reference: tests/golden/tsn_mbv2.npz, heads.npz).  What is checked here is the
orchestration — which entry point gets which operand, shape and saved tensor; the CUDA kernels are checked on the GPU."""
import contextlib
import io

import pytest
import torch
import torch.nn.functional as F

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
    return ehgr_b200


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


@pytest.mark.parametrize("temporal,train_bn", [("tsm", True), ("none", True), ("tsm", False), ("action", True), ("action", False)])
def test_mobilenet_chain_and_head_match_the_oracle(emulated, temporal, train_bn):
    E = emulated
    T, cls = 4, 11
    sd0 = O.build_tsn_state(cls, temporal, 8, seed=2)
    with _quiet():
        model = E.TSN(cls, T, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False,
                      is_shift=(temporal != "none"), shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224,
                      temporal_module=('action' if temporal == "action" else 'tsm'), print_spec=False)
    model.load_state_dict(sd0, strict=True)
    model.train(train_bn)
    for d in model.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    rgb, _, labels = O.synthetic_clip_batch(2, T, 64, cls, seed=4)
    x = rgb.view((-1, 3) + tuple(rgb.shape[-2:]))
    with E.fused.compute_dtype(torch.float32):
        fmap = E.fused.mobilenet_v2_features(model.base_model, x)
        logits = E.fused.classifier_head(model, fmap)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    sd64 = O.clone_state(sd0, dtype=torch.float64)
    ol = O.tsn_forward(rgb.double(), sd64, T, temporal, 8, bn_training=train_bn)
    F.cross_entropy(ol, labels).backward()
    assert rel_err(logits, ol) < 1e-4
    check_grads_up_to_relu_flips(model.named_parameters(), {k: v.grad for k, v in sd64.items() if v.is_floating_point()})
    for k, b in model.named_buffers():
        if b.dtype.is_floating_point:
            assert rel_err(b, sd64[k]) < 1e-4, k
        else:                                   # the functional oracle does not count batches; nn.BatchNorm2d does
            assert int(b) == (1 if train_bn else 0), k


def test_mtmm_wrapper_decoder_and_taps(emulated):
    """models_MTMM.TSN on MobileNetV2: backbone chain -> classifier head + depth decoder (four CONV3 units, materialised
    upsamples, depth head) against the oracle's mtmm_forward; SD-style taps of the chain against the oracle's taps."""
    E = emulated
    T, cls = 2, 7
    sd0 = O.build_mtmm_state(cls, "tsm", 8, seed=3)
    with _quiet():
        model = E.tsn_mtmm.TSN(cls, T, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False,
                               is_shift=True, shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224,
                               modal='rgb_depth', temporal_module='tsm', print_spec=False)
    model.load_state_dict(sd0, strict=True)
    model.train()
    for d in model.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    rgb, depth, labels = O.synthetic_clip_batch(2, T, 64, cls, seed=5)
    x = rgb.view((-1, 3) + tuple(rgb.shape[-2:]))
    with E.fused.compute_dtype(torch.float32):
        fmap = E.fused.mobilenet_v2_features(model.base_model, x)
        logits = E.fused.classifier_head(model, fmap)
        dpred = E.fused.depth_decoder(model.global_decoder, fmap)
    gt = F.interpolate(depth.view(-1, 1, 64, 64), (16, 16), mode='bilinear')
    (F.cross_entropy(logits, labels) + 0.01 * F.mse_loss(dpred, gt)).backward()
    sd64 = O.clone_state(sd0, dtype=torch.float64)
    ol, od = O.mtmm_forward(rgb.double(), sd64, T, "tsm", 8, True)
    (F.cross_entropy(ol, labels) + 0.01 * F.mse_loss(od, gt.double())).backward()
    assert tuple(dpred.shape) == tuple(od.shape) == (2 * T, 1, 16, 16)
    assert rel_err(logits, ol) < 1e-4 and rel_err(dpred, od) < 1e-4
    check_grads_up_to_relu_flips(model.named_parameters(), {k: v.grad for k, v in sd64.items() if v.is_floating_point()})
    # taps (the SD wrapper's features[3], [6], [13] + final map) come out of ONE chain pass
    model.zero_grad()
    with E.fused.compute_dtype(torch.float32):
        taps = E.fused.mobilenet_v2_features(model.base_model, x, taps=(3, 6, 13))
    assert [tuple(t.shape[1:]) for t in taps] == [(24, 16, 16), (32, 8, 8), (96, 4, 4), (1280, 2, 2)]
    sum((t.float() ** 2).mean() for t in taps).backward()
    assert model.base_model.features[0][0].weight.grad.abs().sum() > 0


def _model(E, mod, T, cls, **kw):
    with _quiet():
        m = mod.TSN(cls, T, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False, is_shift=True,
                    shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224, temporal_module='tsm',
                    print_spec=False, **kw)
    m.train()
    for d in m.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    return m


def test_sd_wrapper_and_fused_sd_loss_match_the_oracle(emulated):
    """tsn_sd.TSN.forward (one chain pass with three taps, SepConv exit heads as fused chains, K10 heads) and
    losses.sd_loss (K13: forward + gradients from one launch) against oracle.sd_forward / sd_loss (train_sd.py:227-265)."""
    E = emulated
    T, cls = 2, 6
    sd0 = O.build_sd_state(cls, "tsm", 8, seed=5)
    model = _model(E, E.tsn_sd, T, cls)
    model.load_state_dict(sd0, strict=True)
    rgb, _, labels = O.synthetic_clip_batch(2, T, 64, cls, seed=6)
    with E.fused.compute_dtype(torch.float32):
        outs = model(rgb)
        total, terms = E.losses.sd_loss(outs[:4], outs[4:], labels, 0.1, 1e-6, 3.0)
    total.backward()
    sd64 = O.clone_state(sd0, dtype=torch.float64)
    oouts = O.sd_forward(rgb.double(), sd64, T, "tsm", 8, True)
    ototal, oterms = O.sd_loss(oouts[:4], oouts[4:], labels, 0.1, 1e-6, 3.0)
    ototal.backward()
    for i, (a, b) in enumerate(zip(outs, oouts)):
        assert a.shape == b.shape and rel_err(a, b) < 2e-4, i
    assert abs(total.item() - ototal.item()) < 1e-4 * abs(ototal.item())
    assert rel_err(terms, torch.as_tensor([float(v) for v in oterms["ce"] + oterms["kd"] + oterms["feat"]])) < 1e-3
    check_grads_up_to_relu_flips(model.named_parameters(), {k: v.grad for k, v in sd64.items() if v.is_floating_point()})


def test_mtmm_sd_wrapper_and_combined_loss_match_the_oracle(emulated):
    """tsn_mtmm_sd.TSN.forward (ten outputs, models/models_MTMM_SD.py:431-532) and losses.mtmm_sd_loss
    (train_mtmm_sd.py:240-293) against the oracle's mtmm_sd_forward / mtmm_sd_loss."""
    E = emulated
    T, cls = 2, 6
    sd0 = O.build_mtmm_sd_state(cls, "tsm", 8, seed=5)
    model = _model(E, E.tsn_mtmm_sd, T, cls, modal='rgb_depth')
    model.load_state_dict(sd0, strict=True)
    rgb, depth, labels = O.synthetic_clip_batch(2, T, 64, cls, seed=6)
    with E.fused.compute_dtype(torch.float32):
        outs = model(rgb)
        total, terms, mse = E.losses.mtmm_sd_loss(outs[:4], outs[4:8], outs[9], depth, labels, 0.1, 1e-6, 3.0)
    total.backward()
    sd64 = O.clone_state(sd0, dtype=torch.float64)
    oouts = O.mtmm_sd_forward(rgb.double(), sd64, T, "tsm", 8, True)
    assert len(outs) == len(oouts) == 10
    for i, (a, b) in enumerate(zip(outs, oouts)):
        assert a.shape == b.shape and rel_err(a, b) < 2e-4, i
    # train_mtmm_sd.py:240-293 at this resolution (the oracle's mtmm_sd_loss resizes 224 -> 56)
    gt = F.interpolate(depth.double().view(-1, 1, 64, 64), (16, 16), mode='bilinear')
    osd_total, _ = O.sd_loss(oouts[:4], oouts[4:8], labels, 0.1, 1e-6, 3.0)
    ototal = osd_total + 0.9 * 0.01 * F.mse_loss(oouts[9], gt)
    ototal.backward()
    assert abs(total.item() - ototal.item()) < 1e-4 * abs(ototal.item())
    named = [(k, p) for k, p in model.named_parameters() if not k.startswith("local_decoder.")]   # outs[8] is not in the loss
    check_grads_up_to_relu_flips(named, {k: v.grad for k, v in sd64.items() if v.is_floating_point()})


@pytest.mark.parametrize("c,train_bn", [(32, True), (96, False)])
def test_standalone_action_module_matches_the_oracle(emulated, c, train_bn):
    """Action wrapping a 1x1 convolution, used outside a chain (models/action.py:61-116): the gated tensor is materialised
    (action_ops._ActionGateFunction) and handed to the wrapped module; against oracle.action_forward."""
    E = emulated
    T, h = 4, 6
    rs = __import__("numpy").random.RandomState(3 + c)
    sd0 = {}
    O.action_state(sd0, "m", c, 8, rs)
    O._conv_entry(sd0, "m.net.weight", (2 * c, c, 1, 1), rs)
    with _quiet():
        m = E.Action(torch.nn.Conv2d(c, 2 * c, 1, bias=False), n_segment=T, shift_div=8)
    m.load_state_dict({k[2:]: v for k, v in sd0.items()}, strict=True)
    m.train(train_bn)
    g = torch.Generator().manual_seed(c)
    x = torch.randn(2 * T, c, h, h, generator=g, requires_grad=True)
    gy = torch.randn(2 * T, 2 * c, h, h, generator=g)
    with E.fused.compute_dtype(torch.float32):
        y = m(x)
    y.backward(gy)
    sd64 = O.clone_state(sd0, dtype=torch.float64)
    x64 = x.detach().double().requires_grad_(True)
    y64 = O.action_forward(x64, sd64, "m", T, train_bn)
    y64.backward(gy.double())
    assert rel_err(y, y64) < 1e-5 and rel_err(x.grad, x64.grad) < 1e-4
    for k, p in m.named_parameters():
        ref = sd64["m." + k].grad
        if ref.abs().max() < 1e-12:
            assert p.grad.abs().max() < 1e-6, k
        else:
            assert rel_err(p.grad, ref) < 2e-4, k
