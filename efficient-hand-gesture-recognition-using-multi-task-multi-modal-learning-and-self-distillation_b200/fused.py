"""Execution of the backbone modules on CUDA.

BOOTSTRAP STATE (round 1, first slice): the temporal shift runs on the hand-written kernel; the
convolution / BatchNorm / excitation arithmetic below still goes through the torch CUDA library ops
(cuDNN/cuBLAS) while the fused sm_100a block kernels are brought up one by one.  Nothing here runs
on the CPU: every entry point refuses non-CUDA tensors.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

import contextlib

from . import _lib


@contextlib.contextmanager
def compute_dtype(dtype):
    """Activation storage dtype of the backbone inside the block (fp32 master weights either way)."""
    if dtype == torch.float32:
        yield
    else:
        with torch.autocast("cuda", dtype=dtype):
            yield


def inverted_residual(m, x):
    _lib.require_cuda(x)
    y = m.conv(x)
    return x + y if m.use_res_connect else y


def mobilenet_v2_features(model, x):
    """features[0..18] -> [NT, 1280, H/32, W/32]."""
    _lib.require_cuda(x)
    return model.features(x)


def mobilenet_v2_forward(model, x):
    x = mobilenet_v2_features(model, x)
    x = x.mean(3).mean(2)
    return model.classifier(x)


def classifier_head(tsn, fmap):
    """global average pool -> Dropout -> new_fc -> mean over segments (models/models.py:341-356)."""
    pooled = fmap.mean(3).mean(2)
    drop = getattr(tsn.base_model, tsn.base_model.last_layer_name)
    z = tsn.new_fc(drop(pooled))
    z = z.view((-1, tsn.num_segments) + z.size()[1:])
    return tsn.consensus(z).squeeze(1)


def action_forward(m, x):
    """out = net(x_shift * (3 + g_STE + g_CE + g_ME)) — reference models/action.py:61-116."""
    _lib.require_cuda(x)
    nt, c, h, w = x.shape
    T = m.n_segment
    n = nt // T
    x5 = x.reshape(n, T, c, h, w)
    # per-channel 3-tap temporal FIR (zero padded)
    wt = m.action_shift.weight.view(c, 3)
    xp = F.pad(x5, (0, 0, 0, 0, 0, 0, 1, 1))
    xs = (xp[:, :-2] * wt[:, 0].view(1, 1, c, 1, 1) + xp[:, 1:-1] * wt[:, 1].view(1, 1, c, 1, 1)
          + xp[:, 2:] * wt[:, 2].view(1, 1, c, 1, 1))
    # STE
    g1 = torch.sigmoid(m.action_p1_conv1(xs.mean(2, keepdim=True).transpose(1, 2)))  # [n,1,T,h,w]
    g1 = g1.transpose(1, 2)                                                          # [n,T,1,h,w]
    # CE
    p = xs.mean((3, 4))                                                              # [n,T,c]
    s = F.conv2d(p.reshape(nt, c, 1, 1), m.action_p2_squeeze.weight).view(n, T, -1).transpose(1, 2)
    s = F.relu(m.action_p2_conv1(s)).transpose(1, 2).reshape(nt, -1, 1, 1)
    g2 = torch.sigmoid(F.conv2d(s, m.action_p2_expand.weight)).view(n, T, c, 1, 1)
    # ME
    x3 = m.action_p3_bn1(m.action_p3_squeeze(xs.reshape(nt, c, h, w)))
    cr = x3.shape[1]
    c3 = m.action_p3_conv1(x3).view(n, T, cr, h, w)
    x3 = x3.view(n, T, cr, h, w)
    d = F.pad(c3[:, 1:] - x3[:, :-1], (0, 0, 0, 0, 0, 0, 0, 1))
    g3 = torch.sigmoid(F.conv2d(d.mean((3, 4)).reshape(nt, cr, 1, 1), m.action_p3_expand.weight)).view(n, T, c, 1, 1)
    y = xs * (3.0 + g1 + g2 + g3)
    return m.net(y.reshape(nt, c, h, w))
