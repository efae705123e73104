"""The C-ABI emulator (tests/abi_emulator.py) restates EVERY entry point of include/ehgr_b200.h, and the stand-alone
operators around the chain — TemporalShift, TemporalPool, the uint8 normalisation, the flat SGD + EMA — run on it against
the live-reference fixtures / PyTorch.  Host logic only; the CUDA kernels get the same checks on the GPU."""
import inspect
import re

import numpy as np
import pytest
import torch

import abi_emulator
from conftest import GOLDEN, REPO


@pytest.fixture()
def emulated(monkeypatch):
    import ehgr_b200
    monkeypatch.setattr(ehgr_b200._lib, "call", abi_emulator.call)
    monkeypatch.setattr(ehgr_b200._lib, "require_cuda", lambda *t: None)
    monkeypatch.setattr(ehgr_b200._lib, "stream_ptr", lambda device=None: 0)
    monkeypatch.setattr(ehgr_b200._lib, "on_gpu", lambda t: True)
    monkeypatch.setattr(torch.cuda, "device", lambda *a, **k: __import__("contextlib").nullcontext())
    return ehgr_b200


def test_emulator_restates_every_entry_point_with_the_declared_arity():
    import ehgr_b200
    text = re.sub(r"/\*.*?\*/", "", (REPO / "include" / "ehgr_b200.h").read_text(), flags=re.S)
    declared = set(re.findall(r"\bint\s+(ehgr_[a-z0-9_]+)\s*\(", text)) - {"ehgr_abi_version"}
    assert declared == set(ehgr_b200._lib.SIGNATURES) == set(abi_emulator._TABLE)
    for name, argtypes in ehgr_b200._lib.SIGNATURES.items():
        assert len(inspect.signature(abi_emulator._TABLE[name]).parameters) == len(argtypes), name


def test_temporal_shift_module_bit_exact_against_the_live_reference_fixture(emulated):
    """TemporalShift.shift / its autograd backward (models/temporal_shift.py:27-46) in both layouts."""
    E = emulated
    z = np.load(GOLDEN / "shift.npz")
    for n in sorted({k.split("_")[0] for k in z.files}):
        nt, c, h, w, T, div = (int(v) for v in z[n + "_meta"])
        for fmt in (torch.contiguous_format, torch.channels_last):
            x = torch.from_numpy(z[n + "_x"]).contiguous(memory_format=fmt).requires_grad_(True)
            y = E.TemporalShift.shift(x, T, fold_div=div)
            assert torch.equal(y, torch.from_numpy(z[n + "_y"])), (n, fmt)
            y.backward(torch.from_numpy(z[n + "_g"]).contiguous(memory_format=fmt))
            assert torch.equal(x.grad, torch.from_numpy(z[n + "_gx"])), (n, fmt)
    with pytest.raises(RuntimeError):
        E.TemporalShift.shift(torch.zeros(7, 8, 2, 2), 4, fold_div=8)            # 7 frames are not clips of 4


def test_temporal_pool_bit_exact_against_the_live_reference_fixture(emulated):
    E = emulated
    z = np.load(GOLDEN / "ema_pool.npz")
    for n in ("tp_a", "tp_b", "tp_c"):
        nt, c, h, T = (int(v) for v in z[n + "_meta"])
        x = torch.from_numpy(z[n + "_x"]).requires_grad_(True)
        y = E.TemporalPool.temporal_pool(x, T)
        assert torch.equal(y, torch.from_numpy(z[n + "_y"])), n
        y.backward(torch.ones_like(y) * 0.5 + y.detach())
        assert torch.equal(x.grad, torch.from_numpy(z[n + "_gx"])), n


def test_normalize_u8_is_the_cpu_transform_bit_for_bit(emulated):
    """ToTorchFormatTensor(div=True) + GroupNormalize (models/spatial_transforms.py:66-80,489-503)."""
    E = emulated
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (2, 4, 3, 8, 10), dtype=torch.uint8, generator=g)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    got = E.train_step.normalize_u8(frames, mean, std)
    want = frames.float().div(255)
    want = (want - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1)
    assert torch.equal(got, want)
    depth = torch.randint(0, 256, (2, 4, 1, 8, 10), dtype=torch.uint8, generator=g)
    assert torch.equal(E.train_step.normalize_u8(depth), depth.float().div(255))


def test_flat_sgd_and_ema_follow_torch_sgd_and_the_reference_ema(emulated):
    """FlatSGD = torch.optim.SGD over the nine policy groups (train_mtmm.py:576-585); FlatEMA = EMAWrapper
    (train_mtmm.py:110-140) — three steps with random gradients on a small module tree that has every kind of entry."""
    import contextlib
    import copy
    import io
    E = emulated
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = E.TSN(5, 2, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False, is_shift=True,
                      fc_lr5=True, temporal_module='action', print_spec=False)
    twin = copy.deepcopy(model)
    lr, mom, wd, decay = 0.05, 0.9, 5e-4, 0.9
    ref_opt = E.train_step.build_sgd(twin, lr, mom, wd)
    ema_ref = {k: v.clone() for k, v in twin.state_dict().items()}
    buckets = E.train_step.GradBuckets(list(model.parameters()), 3)
    opt = E.train_step.FlatSGD(model.get_optim_policies(), buckets, lr, mom, wd)
    ema = E.train_step.FlatEMA(model, opt, decay)
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        buckets.zero()
        for p, q in zip(model.parameters(), twin.parameters()):
            gr = torch.randn(p.shape, generator=g)
            p.grad.copy_(gr)
            q.grad = gr.clone()
        for b, bt in zip(model.buffers(), twin.buffers()):          # running statistics move, counters count
            if b.is_floating_point():
                d = torch.randn(b.shape, generator=g) * 0.01
                b.add_(d)
                bt.add_(d)
            else:
                b.add_(1)
                bt.add_(1)
        opt.step(ema.flat, ema.decay)
        ema.update_buffers()
        ref_opt.step()
        # EMAWrapper._update, statement for statement, fed OUR stepped state (torch's SGD may differ from the fused
        # update in the last bit, which is not what the bit-exactness of the EMA expression is about)
        for k, v in model.state_dict().items():
            ema_ref[k].copy_(decay * ema_ref[k] + (1. - decay) * v)
    for (k, p), (_, q) in zip(model.named_parameters(), twin.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-7), k
    for k, v in ema.state_dict().items():
        assert torch.equal(v, ema_ref[k]), k
