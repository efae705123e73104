"""K1 temporal shift: CUDA (through the C ABI) vs the numpy oracle and the reference fixtures.
Bit-exact in every dtype/layout (pure copy)."""
import numpy as np
import pytest
import torch

from oracle import ref_oracle as O
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _shift(x, T, div):
    import ehgr_b200
    return ehgr_b200.TemporalShift.shift(x, T, fold_div=div)


def test_golden_fixtures_fp32_nchw_fwd_bwd():
    z = np.load(GOLDEN / "shift.npz")
    for n in sorted({k.split("_")[0] for k in z.files}):
        nt, c, h, w, T, div = (int(v) for v in z[n + "_meta"])
        x = torch.from_numpy(z[n + "_x"]).cuda().requires_grad_(True)
        y = _shift(x, T, div)
        assert torch.equal(y.cpu(), torch.from_numpy(z[n + "_y"])), n
        y.backward(torch.from_numpy(z[n + "_g"]).cuda())
        assert torch.equal(x.grad.cpu(), torch.from_numpy(z[n + "_gx"])), n


# MobileNetV2 sites (C, H) and a few awkward shapes: odd HW, C not multiple of the vector, fold 0
SHAPES = [(24, 56, 8, 8), (32, 28, 8, 8), (64, 14, 8, 8), (96, 14, 8, 8), (160, 7, 8, 8), (3, 5, 3, 8), (20, 3, 8, 3),
          (9, 3, 2, 2), (8, 1, 1, 4), (50, 7, 4, 5)]


@pytest.mark.parametrize("c,h,T,div", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_matches_oracle_bit_exact(c, h, T, div, dtype, layout):
    rs = np.random.RandomState(c * 131 + h)
    n = 3
    x_np = rs.standard_normal((n * T, c, h, h)).astype(np.float32)
    g_np = rs.standard_normal((n * T, c, h, h)).astype(np.float32)
    x = torch.from_numpy(x_np).to(dtype)
    g = torch.from_numpy(g_np).to(dtype)
    # oracle on the raw bits (uint16 view for bf16): a copy is dtype-agnostic
    view = torch.int16 if dtype == torch.bfloat16 else torch.int32
    want = torch.from_numpy(O.temporal_shift_np(x.view(view).numpy(), T, div)).view(dtype)
    want_g = torch.from_numpy(O.temporal_shift_bwd_np(g.view(view).numpy(), T, div)).view(dtype)
    xc, gc = x.cuda(), g.cuda()
    if layout == "nhwc":
        xc = xc.contiguous(memory_format=torch.channels_last)
        gc = gc.contiguous(memory_format=torch.channels_last)
    xc.requires_grad_(True)
    y = _shift(xc, T, div)
    if layout == "nhwc" and c > 1 and h > 1:
        assert y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y.detach().cpu().view(view), want.view(view))
    y.backward(gc)
    assert torch.equal(xc.grad.cpu().view(view), want_g.view(view))


def test_module_wrapper_and_inplace_flag():
    import contextlib, io
    import ehgr_b200
    with contextlib.redirect_stdout(io.StringIO()):
        m = ehgr_b200.TemporalShift(torch.nn.Identity(), n_segment=8, n_div=8, inplace=True)
    x = torch.randn(16, 24, 7, 7, device="cuda")
    want = torch.from_numpy(O.temporal_shift_np(x.cpu().numpy(), 8, 8))
    assert torch.equal(m(x).cpu(), want)
    # InplaceShift contract (models/temporal_shift.py:49-76)
    x5 = x.clone().view(2, 8, 24, 7, 7)
    out = ehgr_b200.InplaceShift.apply(x5, 3)
    assert out.data_ptr() == x5.data_ptr() and torch.equal(out.view(16, 24, 7, 7).cpu(), want)


def test_empty_and_ragged():
    x = torch.zeros(0, 8, 4, 4, device="cuda")
    assert _shift(x, 8, 8).shape == (0, 8, 4, 4)
    with pytest.raises(RuntimeError):
        _shift(torch.zeros(7, 8, 4, 4, device="cuda"), 4, 8)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_full_size_properties(dtype):
    """BASELINE config #2 size (B=32 -> 256 frames, the 24x56x56 site): adjointness
    <shift(x), g> == <x, shift^T(g)>, element conservation and the boundary-zero count."""
    import ehgr_b200
    T, div, c, h = 8, 8, 24, 56
    g0 = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randint(-8, 9, (256, c, h, h), device="cuda", generator=g0).to(dtype)
    g = torch.randint(-8, 9, (256, c, h, h), device="cuda", generator=g0).to(dtype)
    y = ehgr_b200.temporal_shift(x, T, div)
    gx = ehgr_b200.temporal_shift_module._run_shift(g, T, c // div, backward=True)
    # small integers: the dot products are exact in fp64
    assert (y.double() * g.double()).sum().item() == (x.double() * gx.double()).sum().item()
    fold = c // div
    x5, y5 = x.view(32, T, c, h, h), y.view(32, T, c, h, h)
    assert torch.equal(y5[:, :-1, :fold], x5[:, 1:, :fold]) and torch.equal(y5[:, 1:, fold:2 * fold], x5[:, :-1, fold:2 * fold])
    assert torch.equal(y5[:, :, 2 * fold:], x5[:, :, 2 * fold:])
    assert y5[:, -1, :fold].abs().sum().item() == 0 and y5[:, 0, fold:2 * fold].abs().sum().item() == 0
    # idempotence of the round trip on the interior: shift^T(shift(x)) keeps unshifted channels
    rt = ehgr_b200.temporal_shift_module._run_shift(y, T, fold, backward=True)
    assert torch.equal(rt.view(32, T, c, h, h)[:, :, 2 * fold:], x5[:, :, 2 * fold:])
    assert torch.equal(rt.view(32, T, c, h, h)[:, 1:, :fold], x5[:, 1:, :fold])
