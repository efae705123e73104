"""MobileNetV2 backbone — drop-in for the reference's ``archs/mobilenet_v2.py``.

The module tree is the checkpoint contract: ``features.{i}`` is the stem
(Conv-BN-ReLU6), seventeen ``InvertedResidual`` blocks whose ``conv`` is an indexable
``nn.Sequential`` of 8 children (5 when ``expand_ratio == 1``) and the final 1x1 Conv-BN-ReLU6,
followed by ``classifier``.  ``nn.Conv2d`` / ``nn.BatchNorm2d`` children are parameter containers
with the reference's names; on CUDA the arithmetic is done by the fused sm_100a block kernels in
``fused.py`` (channels-last, BN applied lazily inside the consumer kernel), not by their
``forward``.

Reference: archs/mobilenet_v2.py:7-20 (conv_bn, conv_1x1_bn), :28-66 (InvertedResidual),
:69-129 (MobileNetV2), :132-143 (mobilenet_v2).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

# (expand ratio t, output channels c, repeats n, first stride s) — archs/mobilenet_v2.py:75-84
INVERTED_RESIDUAL_SETTING = (
    (1, 16, 1, 1),
    (6, 24, 2, 2),
    (6, 32, 3, 2),
    (6, 64, 4, 2),
    (6, 96, 3, 1),
    (6, 160, 3, 2),
    (6, 320, 1, 1),
)


def _conv_unit(cin, cout, kernel, stride, groups=1, relu6=True):
    """[Conv2d(bias-free, 'same' padding), BatchNorm2d(, ReLU6)] — the only layer pattern of this network."""
    mods = [nn.Conv2d(cin, cout, kernel, stride, kernel // 2, groups=groups, bias=False), nn.BatchNorm2d(cout)]
    if relu6:
        mods.append(nn.ReLU6(inplace=True))
    return mods


def conv_bn(inp, oup, stride):
    """Stem: dense 3x3 (archs/mobilenet_v2.py:7-12)."""
    return nn.Sequential(*_conv_unit(inp, oup, 3, stride))


def conv_1x1_bn(inp, oup):
    """Last feature layer: pointwise (archs/mobilenet_v2.py:15-20)."""
    return nn.Sequential(*_conv_unit(inp, oup, 1, 1))


def make_divisible(x, divisible_by=8):
    return int(math.ceil(x * 1. / divisible_by) * divisible_by)


def _block_plan(width_mult, first_in=32):
    """(inp, oup, stride, expand_ratio) of the seventeen blocks, in order."""
    cin = first_in
    for t, c, n, s in INVERTED_RESIDUAL_SETTING:
        cout = make_divisible(c * width_mult) if t > 1 else c
        for rep in range(n):
            yield cin, cout, (s if rep == 0 else 1), t
            cin = cout


class InvertedResidual(nn.Module):
    """``conv`` is an indexable Sequential: [pw, bn, relu6,] dw, bn, relu6, pw-linear, bn — 8 children, or 5
    without the expansion (expand_ratio == 1).  TemporalShift / Action wrap ``conv[0]`` of the 8-child form."""

    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        if stride not in (1, 2):
            raise AssertionError(f"stride must be 1 or 2, got {stride}")
        self.stride = stride
        self.use_res_connect = stride == 1 and inp == oup
        hidden_dim = int(inp * expand_ratio)
        chain = _conv_unit(inp, hidden_dim, 1, 1) if expand_ratio != 1 else []
        chain += _conv_unit(hidden_dim, hidden_dim, 3, stride, groups=hidden_dim)
        chain += _conv_unit(hidden_dim, oup, 1, 1, relu6=False)
        self.conv = nn.Sequential(*chain)

    def forward(self, x):
        from . import fused
        return fused.inverted_residual(self, x)


def _init_conv(m):
    fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
    nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan))
    if m.bias is not None:
        nn.init.zeros_(m.bias)


def _init_bn(m):
    nn.init.ones_(m.weight)
    nn.init.zeros_(m.bias)


def _init_linear(m):
    nn.init.normal_(m.weight, 0.0, 0.01)
    nn.init.zeros_(m.bias)


_INITIALISERS = ((nn.Conv2d, _init_conv), (nn.BatchNorm2d, _init_bn), (nn.Linear, _init_linear))


class MobileNetV2(nn.Module):
    def __init__(self, n_class=1000, input_size=224, width_mult=1.):
        super().__init__()
        if input_size % 32:
            raise AssertionError("input_size must be a multiple of 32")
        self.last_channel = make_divisible(1280 * width_mult) if width_mult > 1.0 else 1280
        blocks = [InvertedResidual(*cfg) for cfg in _block_plan(width_mult)]
        tail_in = blocks[-1].conv[-1].num_features
        self.features = nn.Sequential(conv_bn(3, 32, 2), *blocks, conv_1x1_bn(tail_in, self.last_channel))
        self.classifier = nn.Linear(self.last_channel, n_class)
        self._initialize_weights()

    def forward(self, x):
        from . import fused
        return fused.mobilenet_v2_forward(self, x)

    def _initialize_weights(self):
        """He-normal convs, unit BatchNorm, N(0, 0.01) classifier (archs/mobilenet_v2.py:116-129)."""
        for m in self.modules():
            for kind, init in _INITIALISERS:
                if isinstance(m, kind):
                    init(m)
                    break


_IMAGENET_URL = 'https://www.dropbox.com/s/47tyzpofuuyyv1b/mobilenetv2_1.0-f2a8633.pth.tar?dl=1'


def mobilenet_v2(pretrained=True):
    """Same call as the reference (archs/mobilenet_v2.py:132-143): with ``pretrained`` it downloads the ImageNet
    weights from the reference's URL — and fails without network exactly like the reference would."""
    model = MobileNetV2(width_mult=1)
    if pretrained:
        from torch.hub import load_state_dict_from_url
        model.load_state_dict(load_state_dict_from_url(_IMAGENET_URL, progress=True))
    return model
