import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def rel_err(a, b):
    """max|a-b| / max|b| (norm-wise relative error used for fp32 parity: SURVEY §8a tolerances)."""
    import torch
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def check_grads_up_to_relu_flips(named_params, ref_grads):
    """Host-logic tests on the C-ABI emulator (fp32) against an fp64 run of the same network: a ReLU / ReLU6 whose
    pre-activation sits at rounding level (|z| ~ 1e-7) may take the other branch — BatchNorm is applied as raw*scale+shift
    by the kernels' contract and as (x-mean)*invstd*gamma+beta by PyTorch — and ONE such flip in a late layer with a few
    dozen samples per channel moves every upstream gradient by up to a few per cent.  A wrong operand, shape, saved tensor
    or gradient slot is an O(1) error instead.  So: every parameter within 15 % (2-norm) / 50 % (max norm: a flipped element is a whole
    term of a per-channel sum over a few dozen samples) of the reference
    (a flip in the LAST layer shifts all of them, so no tighter bound on a typical parameter holds either; without a flip
    the errors are ~1e-6).  Gradients that vanish analytically (the shift of a BatchNorm that only feeds a
    train-mode BatchNorm) are rounding noise in both runs: errors are measured against at least 1e-4 of the largest gradient
    norm of the model."""
    named_params = list(named_params)
    floor = 1e-4 * max(ref_grads[k].norm().item() for k, _ in named_params)
    e2s = []
    for k, p in named_params:
        assert p.grad is not None, k
        d = p.grad.detach().double() - ref_grads[k].double()
        e2 = d.norm().item() / max(ref_grads[k].norm().item(), floor)
        emax = d.abs().max().item() / max(ref_grads[k].abs().max().item(), floor)
        assert e2 < 0.15 and emax < 0.5, (k, e2, emax)
        e2s.append(e2)
    return max(e2s)
