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


@pytest.mark.parametrize("temporal,train_bn", [("tsm", True), ("none", True), ("tsm", False)])
def test_mobilenet_chain_and_head_match_the_oracle(emulated, temporal, train_bn):
    E = emulated
    T, cls = 4, 11
    sd0 = O.build_tsn_state(cls, temporal, 8, seed=2)
    with _quiet():
        model = E.TSN(cls, T, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False,
                      is_shift=(temporal != "none"), shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224,
                      temporal_module='tsm', print_spec=False)
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
