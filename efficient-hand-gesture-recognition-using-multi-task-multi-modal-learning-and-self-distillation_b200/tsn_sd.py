"""Self-distillation wrapper — drop-in for the reference's ``models/models_SD.py`` ``TSN`` (:104-431)
(and ``models_SD_actionnet.py``), generalised to MobileNetV2.

The reference builds three shallow exit heads on the ResNet stages (``scala1-3`` of ``SepConv`` blocks,
``avgpool1-3``, ``middle_fc1-3``: models/models_SD.py:214-253) and returns
``(output, middle_output1..3, final_fea, middle1_fea..3)`` (:364-431).  For MobileNetV2 (no reference
implementation exists: SURVEY §8a A13) the taps are the outputs of ``features[3]`` (24 ch, 56x56),
``features[6]`` (32 ch, 28x28), ``features[13]`` (96 ch, 14x14) and the final 1280-ch 7x7 map; each
head downsamples with stride-2 ``SepConv`` blocks until it reaches 7x7 with 1280 channels, so the four
feature vectors compared by the feature loss have equal shape, as in the reference (2048 there).

The backbone (and its taps) runs on the fused sm_100a chain, and so do the exit heads: a ``scala`` (a stack of
``SepConv`` blocks) is ONE fused chain of depthwise / pointwise stages with plain-ReLU row operands
(``fused.sepconv_stack``: the kernels of K7 / K8, no library convolution, BatchNorm or activation op), followed by
the pooling and classifier kernels of K10.  ``SepConv.forward`` itself (a stand-alone block) takes the same path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

from .tsn import TSN as _BaseTSN

MBV2_TAPS = (3, 6, 13)            # features indices
MBV2_TAP_CHANNELS = (24, 32, 96)


class SepConv(nn.Module):
    """models/models_SD.py:81-101."""

    def __init__(self, channel_in, channel_out, kernel_size=3, stride=2, padding=1, affine=True):
        super().__init__()
        self.op = nn.Sequential(
            nn.Conv2d(channel_in, channel_in, kernel_size=kernel_size, stride=stride, padding=padding,
                      groups=channel_in, bias=False),
            nn.Conv2d(channel_in, channel_in, kernel_size=1, padding=0, bias=False),
            nn.BatchNorm2d(channel_in, affine=affine),
            nn.ReLU(inplace=False),
            nn.Conv2d(channel_in, channel_in, kernel_size=kernel_size, stride=1, padding=padding,
                      groups=channel_in, bias=False),
            nn.Conv2d(channel_in, channel_out, kernel_size=1, padding=0, bias=False),
            nn.BatchNorm2d(channel_out, affine=affine),
            nn.ReLU(inplace=False),
        )

    def forward(self, x):
        if _lib.on_gpu(x):
            from . import fused
            return fused.sepconv_stack([self], x)
        return self.op(x)               # CPU: plain module arithmetic (shape / policy tests only; no CUDA kernels exist there)


class TSN(_BaseTSN):
    def __init__(self, num_class, num_segments, modality,
                 base_model='resnet101', new_length=None,
                 consensus_type='avg', before_softmax=True,
                 dropout=0.5, img_feature_dim=112,
                 crop_num=1, partial_bn=True, print_spec=True, pretrain='imagenet',
                 is_shift=False, shift_div=8, shift_place='blockres', fc_lr5=False,
                 temporal_pool=False, non_local=False, *, temporal_module='action'):
        if dropout == 0:
            raise ValueError("models_SD.TSN requires dropout > 0 (new_fc is used unconditionally)")
        super().__init__(num_class, num_segments, modality, base_model=base_model, new_length=new_length,
                         consensus_type=consensus_type, before_softmax=before_softmax, dropout=dropout,
                         img_feature_dim=img_feature_dim, crop_num=crop_num, partial_bn=partial_bn,
                         print_spec=print_spec, pretrain=pretrain, is_shift=is_shift, shift_div=shift_div,
                         shift_place=shift_place, fc_lr5=fc_lr5, temporal_pool=temporal_pool,
                         non_local=non_local, temporal_module=temporal_module)
        self._prepare_self_distillation(num_class)

    def _prepare_self_distillation(self, num_classes):
        feat = self.new_fc.in_features
        if self.base_model_name == 'mobilenetv2':
            c1, c2, c3 = MBV2_TAP_CHANNELS
        else:                                   # ResNet bottleneck stages (models/models_SD.py:215-253)
            c1, c2, c3 = 256, 512, 1024
        self.scala1 = nn.Sequential(SepConv(c1, c2), SepConv(c2, c3), SepConv(c3, feat))
        self.avgpool1 = nn.AdaptiveAvgPool2d((1, 1))
        self.middle_fc1 = nn.Linear(feat, num_classes)
        self.scala2 = nn.Sequential(SepConv(c2, c3), SepConv(c3, feat))
        self.avgpool2 = nn.AdaptiveAvgPool2d((1, 1))
        self.middle_fc2 = nn.Linear(feat, num_classes)
        self.scala3 = nn.Sequential(SepConv(c3, feat))
        self.avgpool3 = nn.AdaptiveAvgPool2d((1, 1))
        self.middle_fc3 = nn.Linear(feat, num_classes)

    def _taps(self, x):
        from . import fused
        bm = self.base_model
        if self.base_model_name == 'mobilenetv2':
            return fused.mobilenet_v2_features(bm, x, taps=MBV2_TAPS)          # (f3, f6, f13, final)
        if _lib.on_gpu(x) and self._fused_resnet():                 # N3: Bottleneck ResNet on the library's kernels
            from . import resnet_ops
            return resnet_ops.resnet_features(bm, x, taps=(1, 2, 3))          # (layer1, layer2, layer3, layer4)
        x = bm.maxpool(bm.relu(bm.bn1(bm.conv1(x))))
        t1 = bm.layer1(x)
        t2 = bm.layer2(t1)
        t3 = bm.layer3(t2)
        return t1, t2, t3, bm.layer4(t3)

    def _exit(self, x, scala, pool, fc):
        from . import fused
        if _lib.on_gpu(x) and all(isinstance(m, SepConv) for m in scala):
            y = fused.sepconv_stack(list(scala), x)     # one fused chain per exit head
            pooled = fused.global_avg_pool(y)           # [NT, F] fp32
            fea = pooled.view(pooled.shape[0], pooled.shape[1], 1, 1)
            return fused.fc_consensus(pooled, fc, self.num_segments), fea
        y = scala(x)
        fea = pool(y)                                   # [NT, F, 1, 1]
        z = fused.fc_consensus(torch.flatten(fea, 1).float(), fc, self.num_segments)
        return z, fea

    def forward(self, x):
        from . import fused
        x = x.view((-1, 3 * self.new_length) + x.size()[-2:])
        t1, t2, t3, fmap = self._taps(x)
        m1, f1 = self._exit(t1, self.scala1, self.avgpool1, self.middle_fc1)
        m2, f2 = self._exit(t2, self.scala2, self.avgpool2, self.middle_fc2)
        m3, f3 = self._exit(t3, self.scala3, self.avgpool3, self.middle_fc3)
        pooled = fused.global_avg_pool(fmap)            # [NT, F] fp32
        final_fea = pooled.view(pooled.shape[0], pooled.shape[1], 1, 1)
        drop = getattr(self.base_model, self.base_model.last_layer_name)
        output = fused.fc_consensus(drop(pooled), self.new_fc, self.num_segments)
        return output, m1, m2, m3, final_fea, f1, f2, f3
