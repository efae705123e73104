"""TSN wrapper — drop-in for the reference's ``models/models.py`` ``TSN`` on the RGB path.

Reference: constructor/flags models/models.py:13-104, backbone preparation :107-212 (MobileNetV2
branch :169-194), BN freezing ``train()`` :214-230, ``get_optim_policies`` :235-321, ``forward``
:323-356.  Same constructor signature and attributes (``base_model``, ``new_fc``, ``consensus``,
``num_segments``, ``base_model.last_layer_name`` ...), same parameter names.

Additions (keyword-only, after the reference's arguments):
  * ``temporal_module='action' | 'tsm'`` — with ``is_shift=True`` the reference inserts ``Action``
    into MobileNetV2 (models/models.py:180-185); ``'tsm'`` inserts ``TemporalShift`` at the same ten
    sites (the "TSM-MobileNetV2" of BASELINE.json).
Only ``modality='RGB'`` is on the hot path; Flow / RGBDiff raise NotImplementedError.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from torch.nn.init import constant_, normal_

from .basic_ops import ConsensusModule


class TSN(nn.Module):
    def __init__(self, num_class, num_segments, modality,
                 base_model='resnet101', new_length=None,
                 consensus_type='avg', before_softmax=True,
                 dropout=0.5, img_feature_dim=112,
                 crop_num=1, partial_bn=True, print_spec=True, pretrain='imagenet',
                 is_shift=False, shift_div=8, shift_place='blockres', fc_lr5=False,
                 temporal_pool=False, non_local=False, *, temporal_module='action'):
        super().__init__()
        if modality != 'RGB':
            raise NotImplementedError("only modality='RGB' is on the B200 hot path")
        if not before_softmax and consensus_type != 'avg':
            raise ValueError("Only avg consensus can be used after Softmax")
        if non_local:
            raise NotImplementedError("non_local blocks are not part of this path")
        if temporal_module not in ('action', 'tsm'):
            raise ValueError("temporal_module must be 'action' or 'tsm'")
        self.modality = modality
        self.num_segments = num_segments
        self.reshape = True
        self.before_softmax = before_softmax
        self.dropout = dropout
        self.crop_num = crop_num
        self.consensus_type = consensus_type
        self.img_feature_dim = img_feature_dim
        self.pretrain = pretrain
        self.is_shift = is_shift
        self.shift_div = shift_div
        self.shift_place = shift_place
        self.base_model_name = base_model
        self.fc_lr5 = fc_lr5
        self.temporal_pool = temporal_pool
        self.non_local = non_local
        self.temporal_module = temporal_module
        self.new_length = 1 if new_length is None else new_length
        if print_spec:
            print(("""
    Initializing TSN with base model: {}.
    TSN Configurations:
        input_modality:     {}
        num_segments:       {}
        new_length:         {}
        consensus_module:   {}
        dropout_ratio:      {}
        img_feature_dim:    {}
            """.format(base_model, self.modality, self.num_segments, self.new_length, consensus_type,
                       self.dropout, self.img_feature_dim)))

        self._prepare_base_model(base_model)
        self._prepare_tsn(num_class)
        self.consensus = ConsensusModule(consensus_type)
        if not self.before_softmax:
            self.softmax = nn.Softmax()
        self._enable_pbn = partial_bn
        if partial_bn:
            self.partialBN(True)

    # ---- construction ---------------------------------------------------------------------
    def _prepare_tsn(self, num_class):
        name = self.base_model.last_layer_name
        feature_dim = getattr(self.base_model, name).in_features
        std = 0.001
        if self.dropout == 0:
            head = nn.Linear(feature_dim, num_class)
            setattr(self.base_model, name, head)
            self.new_fc = None
        else:
            setattr(self.base_model, name, nn.Dropout(p=self.dropout))
            head = self.new_fc = nn.Linear(feature_dim, num_class)
        normal_(head.weight, 0, std)
        constant_(head.bias, 0)
        return feature_dim

    def _insert_temporal(self, net):
        if self.temporal_module == 'tsm':
            from .temporal_shift import make_temporal_shift
        else:
            from .action import make_temporal_shift
        make_temporal_shift(net, self.num_segments, n_div=self.shift_div, place=self.shift_place,
                            temporal_pool=self.temporal_pool)

    def _prepare_base_model(self, base_model):
        print('=> base model: {}'.format(base_model))
        pretrained = self.pretrain == 'imagenet'
        if 'resnet' in base_model:
            import torchvision
            self.base_model = getattr(torchvision.models, base_model)(pretrained)
            if self.is_shift:
                print('Adding action...')
                self._insert_temporal(self.base_model)
            self.base_model.last_layer_name = 'fc'
        elif base_model == 'mobilenetv2':
            from .mobilenet_v2 import mobilenet_v2
            self.base_model = mobilenet_v2(pretrained)
            self.base_model.last_layer_name = 'classifier'
            if self.is_shift:
                self._insert_temporal(self.base_model)
        else:
            raise ValueError('Unknown base model: {}'.format(base_model))
        self.input_size = 224
        self.input_mean = [0.485, 0.456, 0.406]
        self.input_std = [0.229, 0.224, 0.225]
        self.base_model.avgpool = nn.AdaptiveAvgPool2d(1)

    # ---- BN freezing (models/models.py:214-233) -------------------------------------------
    def train(self, mode=True):
        super().train(mode)
        if self._enable_pbn and mode:
            print("Freezing BatchNorm2D except the first one.")
            count = 0
            for m in self.base_model.modules():
                if isinstance(m, nn.BatchNorm2d):
                    count += 1
                    if count >= 2:
                        m.eval()
                        m.weight.requires_grad = False
                        m.bias.requires_grad = False
        return self

    def partialBN(self, enable):
        self._enable_pbn = enable

    # ---- optimiser groups (models/models.py:235-321) ---------------------------------------
    def get_optim_policies(self):
        groups = {k: [] for k in ("first_conv_weight", "first_conv_bias", "normal_weight", "normal_bias",
                                  "bn", "custom_weight", "custom_bn", "lr5_weight", "lr10_bias")}
        conv_cnt = bn_cnt = 0
        # ConvTranspose2d: the MTMM+SD decoders (models/models_MTMM_SD.py:361 lists it with the convolutions)
        conv_types = (nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d)
        bn_types = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)
        for name, m in self.named_modules():
            if 'action' in name:
                ps = list(m.parameters())
                if 'bn' not in name:
                    groups["custom_weight"].append(ps[0])
                elif not self._enable_pbn or bn_cnt == 1:
                    groups["custom_bn"].extend(ps)
            elif isinstance(m, conv_types):
                ps = list(m.parameters())
                conv_cnt += 1
                w, b = ("first_conv_weight", "first_conv_bias") if conv_cnt == 1 else ("normal_weight", "normal_bias")
                groups[w].append(ps[0])
                if len(ps) == 2:
                    groups[b].append(ps[1])
            elif isinstance(m, nn.Linear):
                ps = list(m.parameters())
                groups["lr5_weight" if self.fc_lr5 else "normal_weight"].append(ps[0])
                if len(ps) == 2:
                    groups["lr10_bias" if self.fc_lr5 else "normal_bias"].append(ps[1])
            elif isinstance(m, bn_types):
                bn_cnt += 1
                if not self._enable_pbn or bn_cnt == 1:
                    groups["bn"].extend(list(m.parameters()))
            elif len(m._modules) == 0 and len(list(m.parameters())) > 0:
                raise ValueError("New atomic module type: {}. Need to give it a learning policy".format(type(m)))
        spec = (("first_conv_weight", 1, 1, "first_conv_weight"), ("first_conv_bias", 2, 0, "first_conv_bias"),
                ("normal_weight", 1, 1, "normal_weight"), ("normal_bias", 2, 0, "normal_bias"),
                ("bn", 1, 0, "BN scale/shift"), ("custom_weight", 1, 1, "custom_weight"),
                ("custom_bn", 1, 0, "custom_bn"), ("lr5_weight", 5, 1, "lr5_weight"),
                ("lr10_bias", 10, 0, "lr10_bias"))
        return [{'params': groups[k], 'lr_mult': lr, 'decay_mult': dm, 'name': nm} for k, lr, dm, nm in spec]

    # ---- forward (models/models.py:323-356) ------------------------------------------------
    def forward(self, input, no_reshape=False):
        assert input.size()[1] > 3, \
            'channel and temporal dimension mismatch, tensor size should be: n_batch, n_segment, nc, h, w'
        if (_lib.on_gpu(input) and not no_reshape and self.reshape and self.base_model_name == 'mobilenetv2'
                and not (self.is_shift and self.temporal_pool)):
            # backbone -> global average pool -> Dropout -> new_fc -> segment consensus, all on the library's kernels
            # (csrc/head.cu folds the 'avg' consensus in front of the classifier GEMV: fused.classifier_head)
            from . import fused
            fmap = fused.mobilenet_v2_features(self.base_model, input.view((-1, 3 * self.new_length) + input.size()[-2:]))
            return fused.classifier_head(self, fmap)
        if (_lib.on_gpu(input) and not no_reshape and self.reshape and not (self.is_shift and self.temporal_pool)
                and self._fused_resnet()):
            # N3: torchvision Bottleneck ResNet (+ TemporalShift on conv1) on the library's kernels (resnet_ops.py)
            from . import fused, resnet_ops
            fmap = resnet_ops.resnet_features(self.base_model, input.view((-1, 3 * self.new_length) + input.size()[-2:]))
            return fused.classifier_head(self, fmap)
        if not no_reshape:
            sample_len = 3 * self.new_length
            base_out = self.base_model(input.view((-1, sample_len) + input.size()[-2:]))
        else:
            base_out = self.base_model(input)
        if self.dropout > 0:
            base_out = self.new_fc(base_out)
        if not self.before_softmax:
            base_out = self.softmax(base_out)
        if self.reshape:
            segs = self.num_segments // 2 if (self.is_shift and self.temporal_pool) else self.num_segments
            base_out = base_out.view((-1, segs) + base_out.size()[1:])
            output = self.consensus(base_out)
            return output.squeeze(1)

    def _fused_resnet(self) -> bool:
        """True when the backbone is a torchvision ResNet the sm_100a path covers (resnet_ops.plan_of).  Anything else
        (BasicBlock nets, Action on wide stages, temporal_pool ...) runs the torchvision modules, and says so once."""
        if 'resnet' not in self.base_model_name:
            return False
        from . import resnet_ops
        ok, why = resnet_ops.supported(self.base_model)
        if not ok and not getattr(self, '_warned_library_resnet', False):
            import warnings
            warnings.warn("ResNet backbone runs on torchvision / library kernels, not on libehgr_b200: " + why)
            self._warned_library_resnet = True
        return ok

    @property
    def crop_size(self):
        return self.input_size

    @property
    def scale_size(self):
        return self.input_size * 256 // 224
