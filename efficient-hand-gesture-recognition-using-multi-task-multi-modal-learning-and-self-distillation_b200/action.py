"""ACTION module (spatio-temporal / channel / motion excitation) — drop-in for the reference's
``models/action.py`` (``Action`` :8-116, ``make_temporal_shift`` :179-233; the file's duplicate
``TemporalShift`` / ``TemporalPool`` :119-176 are re-exported from ``temporal_shift``).

Parameter containers keep the reference's names (``action_shift``, ``action_p1_conv1``,
``action_p2_squeeze/conv1/expand``, ``action_p3_squeeze/bn1/conv1/expand``) because checkpoints
and ``get_optim_policies`` key on them (models/models.py:249-257).  The arithmetic uses the
identity  out = net( x_shift * (3 + g_STE + g_CE + g_ME) )   and runs in the
fused kernels of ``fused.py``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .temporal_shift import TemporalPool, TemporalShift  # noqa: F401  (same names as the reference file)


class Action(nn.Module):
    def __init__(self, net, n_segment=3, shift_div=8):
        super().__init__()
        self.net = net
        self.n_segment = n_segment
        self.in_channels = self.net.in_channels
        self.out_channels = self.net.out_channels
        self.kernel_size = self.net.kernel_size
        self.stride = self.net.stride
        self.padding = self.net.padding
        self.reduced_channels = self.in_channels // 16
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.relu = nn.ReLU(inplace=True)
        self.sigmoid = nn.Sigmoid()
        self.fold = self.in_channels // shift_div
        C, Cr = self.in_channels, self.reduced_channels

        # learnable per-channel 3-tap temporal filter, initialised to the TSM shift pattern
        self.action_shift = nn.Conv1d(C, C, kernel_size=3, padding=1, groups=C, bias=False)
        w = torch.zeros_like(self.action_shift.weight)
        w[:self.fold, 0, 2] = 1                 # frame t takes t+1
        w[self.fold:2 * self.fold, 0, 0] = 1    # frame t takes t-1
        if 2 * self.fold < C:
            w[2 * self.fold:, 0, 1] = 1         # untouched channels
        self.action_shift.weight.data.copy_(w)

        # spatio-temporal excitation
        self.action_p1_conv1 = nn.Conv3d(1, 1, kernel_size=(3, 3, 3), stride=(1, 1, 1), bias=False,
                                         padding=(1, 1, 1))
        # channel excitation
        self.action_p2_squeeze = nn.Conv2d(C, Cr, kernel_size=(1, 1), stride=(1, 1), bias=False, padding=(0, 0))
        self.action_p2_conv1 = nn.Conv1d(Cr, Cr, kernel_size=3, stride=1, bias=False, padding=1, groups=1)
        self.action_p2_expand = nn.Conv2d(Cr, C, kernel_size=(1, 1), stride=(1, 1), bias=False, padding=(0, 0))
        # motion excitation
        self.pad = (0, 0, 0, 0, 0, 0, 0, 1)
        self.action_p3_squeeze = nn.Conv2d(C, Cr, kernel_size=(1, 1), stride=(1, 1), bias=False, padding=(0, 0))
        self.action_p3_bn1 = nn.BatchNorm2d(Cr)
        self.action_p3_conv1 = nn.Conv2d(Cr, Cr, kernel_size=(3, 3), stride=(1, 1), bias=False, padding=(1, 1),
                                         groups=Cr)
        self.action_p3_expand = nn.Conv2d(Cr, C, kernel_size=(1, 1), stride=(1, 1), bias=False, padding=(0, 0))
        print('=> Using ACTION')

    def forward(self, x):
        from . import fused
        return fused.action_forward(self, x)


def make_temporal_shift(net, n_segment, n_div=8, place='blockres', temporal_pool=False):
    """Insert ``Action`` into a backbone — reference models/action.py:179-233 (ResNet ``conv1`` of
    every bottleneck), extended to MobileNetV2 with the predicate of models/models.py:183."""
    if temporal_pool:
        n_segment_list = [n_segment, n_segment // 2, n_segment // 2, n_segment // 2]
    else:
        n_segment_list = [n_segment] * 4
    assert n_segment_list[-1] > 0
    print('=> n_segment per stage: {}'.format(n_segment_list))

    import torchvision
    from .temporal_shift import _is_mobilenet_v2, residual_sites
    if isinstance(net, torchvision.models.ResNet):
        if 'blockres' not in place:
            # the reference's 'block' branch is a pdb breakpoint (models/action.py:202)
            raise NotImplementedError(place)
        n_round = 1
        if len(list(net.layer3.children())) >= 23:
            n_round = 2
            print('=> Using n_round {} to insert temporal shift'.format(n_round))
        for name, seg in zip(['layer1', 'layer2', 'layer3', 'layer4'], n_segment_list):
            blocks = list(getattr(net, name).children())
            print('=> Processing stage with {} blocks residual'.format(len(blocks)))
            for i, b in enumerate(blocks):
                if i % n_round == 0:
                    b.conv1 = Action(b.conv1, n_segment=seg, shift_div=n_div)
            setattr(net, name, nn.Sequential(*blocks))
    elif _is_mobilenet_v2(net):
        for m in residual_sites(net):
            m.conv[0] = Action(m.conv[0], n_segment=n_segment, shift_div=n_div)
    else:
        raise NotImplementedError(place)
