"""MTMM + self-distillation wrapper — drop-in for the reference's ``models/models_MTMM_SD.py`` ``TSN`` (:104-532),
generalised to MobileNetV2.

The reference (ResNet only) adds to the self-distillation network (three ``SepConv`` exit heads, :274-313) two
depth decoders made of ``ConvTranspose2d(k=4, s=2, p=1)`` layers (:226-249): ``local_decoder`` on the ``maxpool`` output
(64 ch, 56x56 -> [NT,1,224,224]) and ``global_decoder`` on ``layer4`` (2048 ch, 7x7 -> [NT,1,56,56]); it runs the
backbone twice (by hand :431-476 and through a torch.fx feature extractor :492) and returns, for ``modal='rgb_depth'``,
ten tensors (:522-523).  Here the backbone runs ONCE on the fused chain with four taps; on MobileNetV2 (no reference
implementation: SURVEY §8a) ``local_decoder`` reads the ``features[3]`` output (24 ch, 56x56 — the first exit head's
tap) and ``global_decoder`` the final 1280-channel 7x7 map.  One consequence of the single pass: BatchNorm running
statistics are updated once per step, not twice with the same batch statistics as in the reference.

The skeleton / text modalities (:251-272) need annotation files that do not ship with the reference; they are not built.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .tsn_sd import MBV2_TAP_CHANNELS, TSN as _SDTSN


def make_convt_decoder(chans) -> nn.Sequential:
    """ConvTranspose2d(k4, s2, p1) [+ BatchNorm2d] ... ConvTranspose2d, Sigmoid — models/models_MTMM_SD.py:227-249."""
    layers = []
    for j, (ci, co) in enumerate(zip(chans[:-1], chans[1:])):
        layers.append(nn.ConvTranspose2d(ci, co, kernel_size=4, stride=2, padding=1))
        if j + 2 < len(chans):
            layers.append(nn.BatchNorm2d(co))
    layers.append(nn.Sigmoid())
    return nn.Sequential(*layers)


class TSN(_SDTSN):
    def __init__(self, num_class, num_segments, modality,
                 base_model='resnet101', new_length=None,
                 consensus_type='avg', before_softmax=True,
                 dropout=0.5, img_feature_dim=112,
                 crop_num=1, partial_bn=True, print_spec=True, pretrain='imagenet',
                 is_shift=False, shift_div=8, shift_place='blockres', fc_lr5=False,
                 temporal_pool=False, non_local=False,
                 modal='rgb_depth', *, temporal_module='action'):
        if modal not in ('rgb', 'rgb_depth'):
            raise NotImplementedError(f"modal={modal!r}: the skeleton / text branches need data the reference does not ship")
        self.modal = modal
        super().__init__(num_class, num_segments, modality, base_model=base_model, new_length=new_length,
                         consensus_type=consensus_type, before_softmax=before_softmax, dropout=dropout,
                         img_feature_dim=img_feature_dim, crop_num=crop_num, partial_bn=partial_bn,
                         print_spec=print_spec, pretrain=pretrain, is_shift=is_shift, shift_div=shift_div,
                         shift_place=shift_place, fc_lr5=fc_lr5, temporal_pool=temporal_pool,
                         non_local=non_local, temporal_module=temporal_module)
        if self.modal.find('depth') != -1:
            feat = self.new_fc.in_features
            local_in = MBV2_TAP_CHANNELS[0] if self.base_model_name == 'mobilenetv2' else 64
            self.local_decoder = make_convt_decoder((local_in, 32, 1))
            self.global_decoder = make_convt_decoder((feat, 256, 32, 1))

    def forward(self, x):
        from . import fused
        x = x.view((-1, 3 * self.new_length) + x.size()[-2:])
        if self.base_model_name == 'mobilenetv2':
            t1, t2, t3, fmap = self._taps(x)
            local_in = t1
        elif _lib.on_gpu(x) and self._fused_resnet():      # N3: one pass of resnet_ops with the max-pool output tapped too
            from . import resnet_ops
            local_in, t1, t2, t3, fmap = resnet_ops.resnet_features(self.base_model, x, taps=(0, 1, 2, 3))
        else:
            bm = self.base_model
            local_in = bm.maxpool(bm.relu(bm.bn1(bm.conv1(x))))
            t1 = bm.layer1(local_in)
            t2 = bm.layer2(t1)
            t3 = bm.layer3(t2)
            fmap = bm.layer4(t3)
        m1, f1 = self._exit(t1, self.scala1, self.avgpool1, self.middle_fc1)
        m2, f2 = self._exit(t2, self.scala2, self.avgpool2, self.middle_fc2)
        m3, f3 = self._exit(t3, self.scala3, self.avgpool3, self.middle_fc3)
        pooled = fused.global_avg_pool(fmap)
        final_fea = pooled.view(pooled.shape[0], pooled.shape[1], 1, 1)
        drop = getattr(self.base_model, self.base_model.last_layer_name)
        output = fused.fc_consensus(drop(pooled), self.new_fc, self.num_segments)
        outs = (output, m1, m2, m3, final_fea, f1, f2, f3)
        if self.modal == 'rgb':
            return outs
        return outs + (self.local_decoder(local_in), self.global_decoder(fmap))
