"""MTMM wrapper (classification + auxiliary depth regression) — drop-in for the reference's
``models/models_MTMM.py`` ``TSN`` (:17-299), generalised to MobileNetV2.

The reference accepts torchvision ResNets only (models/models_MTMM.py:112-157) and taps
``layer4`` through a torch.fx feature extractor (:70-77).  Here the backbone's last feature map is
taken by hand (MobileNetV2: ``features[18]`` output, [NT,1280,7,7]; ResNet: ``layer4``), and
``global_decoder`` (:129-155) keeps its architecture with the first convolution's in-channels equal
to the backbone's feature width (1280 for MobileNetV2, 2048 for ResNet-50).

forward(input[N,T,3,H,W]) -> logits[N,cls]                      (modal='rgb')
                           -> (logits[N,cls], depth[NT,1,56,56]) (modal='rgb_depth')
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

from .tsn import TSN as _BaseTSN


def make_global_decoder(in_channels: int) -> nn.Sequential:
    """models/models_MTMM.py:129-155."""
    def unit(i, o, up):
        layers = [nn.Conv2d(i, o, kernel_size=3, stride=1, padding=1, bias=False), nn.BatchNorm2d(o),
                  nn.ReLU(inplace=True)]
        if up:
            layers.append(nn.Upsample(scale_factor=2, mode='nearest'))
        return layers
    return nn.Sequential(*unit(in_channels, 256, True), *unit(256, 64, True), *unit(64, 32, True),
                         *unit(32, 32, False), nn.Conv2d(32, 1, kernel_size=1, stride=1, padding=0), nn.Sigmoid())


class TSN(_BaseTSN):
    def __init__(self, num_class, num_segments, modality,
                 base_model='resnet101', new_length=None,
                 consensus_type='avg', before_softmax=True,
                 dropout=0.5, img_feature_dim=112,
                 crop_num=1, partial_bn=True, print_spec=True, pretrain='imagenet',
                 is_shift=False, shift_div=8, shift_place='blockres', fc_lr5=False,
                 temporal_pool=False, non_local=False,
                 modal='rgb_depth', *, temporal_module='action'):
        if dropout == 0:
            # the reference calls self.new_fc unconditionally (models/models_MTMM.py:280)
            raise ValueError("models_MTMM.TSN requires dropout > 0 (new_fc is used unconditionally)")
        self.modal = modal
        super().__init__(num_class, num_segments, modality, base_model=base_model, new_length=new_length,
                         consensus_type=consensus_type, before_softmax=before_softmax, dropout=dropout,
                         img_feature_dim=img_feature_dim, crop_num=crop_num, partial_bn=partial_bn,
                         print_spec=print_spec, pretrain=pretrain, is_shift=is_shift, shift_div=shift_div,
                         shift_place=shift_place, fc_lr5=fc_lr5, temporal_pool=temporal_pool,
                         non_local=non_local, temporal_module=temporal_module)
        if self.modal.find('depth') != -1:
            self.global_decoder = make_global_decoder(self.new_fc.in_features)

    def _last_feature_map(self, x):
        from . import fused
        bm = self.base_model
        if self.base_model_name == 'mobilenetv2':
            return fused.mobilenet_v2_features(bm, x)
        if _lib.on_gpu(x) and self._fused_resnet():                 # N3: Bottleneck ResNet on the library's kernels
            from . import resnet_ops
            return resnet_ops.resnet_features(bm, x)
        x = bm.maxpool(bm.relu(bm.bn1(bm.conv1(x))))
        return bm.layer4(bm.layer3(bm.layer2(bm.layer1(x))))

    def forward(self, input):
        from . import fused
        x = input.view((-1, 3 * self.new_length) + input.size()[-2:])
        fmap = self._last_feature_map(x)                       # [NT, F, 7, 7]
        output = fused.classifier_head(self, fmap)             # pool -> dropout -> new_fc -> consensus
        if self.modal == 'rgb':
            return output
        if self.modal == 'rgb_depth':
            if _lib.on_gpu(fmap):                                   # four implicit-GEMM conv stages + the depth head (fused.py)
                return output, fused.depth_decoder(self.global_decoder, fmap)
            return output, self.global_decoder(fmap)           # CPU: shape / policy tests only
        raise ValueError(self.modal)
