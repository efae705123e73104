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


def conv_bn(inp, oup, stride):
    return nn.Sequential(
        nn.Conv2d(inp, oup, 3, stride, 1, bias=False),
        nn.BatchNorm2d(oup),
        nn.ReLU6(inplace=True),
    )


def conv_1x1_bn(inp, oup):
    return nn.Sequential(
        nn.Conv2d(inp, oup, 1, 1, 0, bias=False),
        nn.BatchNorm2d(oup),
        nn.ReLU6(inplace=True),
    )


def make_divisible(x, divisible_by=8):
    return int(math.ceil(x * 1. / divisible_by) * divisible_by)


class InvertedResidual(nn.Module):
    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        assert stride in [1, 2]
        self.stride = stride
        hidden_dim = int(inp * expand_ratio)
        self.use_res_connect = self.stride == 1 and inp == oup

        layers = []
        if expand_ratio != 1:
            layers += [nn.Conv2d(inp, hidden_dim, 1, 1, 0, bias=False),          # pw
                       nn.BatchNorm2d(hidden_dim), nn.ReLU6(inplace=True)]
        layers += [nn.Conv2d(hidden_dim, hidden_dim, 3, stride, 1, groups=hidden_dim, bias=False),  # dw
                   nn.BatchNorm2d(hidden_dim), nn.ReLU6(inplace=True),
                   nn.Conv2d(hidden_dim, oup, 1, 1, 0, bias=False),              # pw-linear
                   nn.BatchNorm2d(oup)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        from . import fused
        return fused.inverted_residual(self, x)


class MobileNetV2(nn.Module):
    def __init__(self, n_class=1000, input_size=224, width_mult=1.):
        super().__init__()
        input_channel = 32
        last_channel = 1280
        assert input_size % 32 == 0
        self.last_channel = make_divisible(last_channel * width_mult) if width_mult > 1.0 else last_channel
        features = [conv_bn(3, input_channel, 2)]
        for t, c, n, s in INVERTED_RESIDUAL_SETTING:
            output_channel = make_divisible(c * width_mult) if t > 1 else c
            for i in range(n):
                features.append(InvertedResidual(input_channel, output_channel, s if i == 0 else 1,
                                                 expand_ratio=t))
                input_channel = output_channel
        features.append(conv_1x1_bn(input_channel, self.last_channel))
        self.features = nn.Sequential(*features)
        self.classifier = nn.Linear(self.last_channel, n_class)
        self._initialize_weights()

    def forward(self, x):
        from . import fused
        return fused.mobilenet_v2_forward(self, x)

    def _initialize_weights(self):
        # archs/mobilenet_v2.py:116-129
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()


def mobilenet_v2(pretrained=True):
    model = MobileNetV2(width_mult=1)
    if pretrained:
        # The reference downloads ImageNet weights from a Dropbox URL (archs/mobilenet_v2.py:135-142).
        # Honour the same call; it fails without network exactly like the reference would.
        from torch.hub import load_state_dict_from_url
        state_dict = load_state_dict_from_url(
            'https://www.dropbox.com/s/47tyzpofuuyyv1b/mobilenetv2_1.0-f2a8633.pth.tar?dl=1', progress=True)
        model.load_state_dict(state_dict)
    return model
