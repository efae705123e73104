"""Loss heads of the two training stages.

* ``mtmm_loss``  — train_mtmm.py:223-231: CE(logits, labels) + 0.01 * MSE(depth_pred,
  bilinear56(depth_gt)).  The 224->56 bilinear resize with align_corners=False is exactly the mean of
  the 2x2 pixels at rows/cols 4i+1, 4i+2, which is what the fused kernel reads.
* ``sd_loss``    — train_sd.py:178-193,227-265: 4 CE + 3 temperature-KL (x T^2) + 3 masked-L2 feature
  terms, weights (1-alpha), alpha, beta.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib


def mtmm_loss(logits, labels, depth_pred, depth_gt, depth_weight: float = 0.01):
    """Returns (loss, depth_mse).  depth_gt: [N,T,1,224,224] (any leading shape, HxW = 4x depth_pred)."""
    _lib.require_cuda(logits, depth_pred, depth_gt)
    gt = depth_gt.reshape(-1, 1, depth_gt.size(-2), depth_gt.size(-1)).float()
    gt = F.interpolate(gt, size=tuple(depth_pred.shape[-2:]), mode='bilinear')
    depth_mse = F.mse_loss(depth_pred.float(), gt)
    return F.cross_entropy(logits.float(), labels) + depth_weight * depth_mse, depth_mse


def sd_loss(outputs, feats, labels, alpha: float = 0.1, beta: float = 1e-6, temperature: float = 3.0):
    """outputs = (final, mid1, mid2, mid3) logits [N,cls]; feats = (final, mid1, mid2, mid3) pooled
    features.  Returns (total, terms[10]) with terms = 4 CE, 3 KD (already x T^2), 3 feature sums."""
    _lib.require_cuda(*outputs, *feats)
    out = [o.float() for o in outputs]
    ce = [F.cross_entropy(o, labels) for o in out]
    soft = torch.softmax(out[0] / temperature, dim=1).detach()
    kd = [-(torch.log_softmax(o / temperature, dim=1) * soft).sum(1).mean() * temperature ** 2 for o in out[1:]]
    f4 = feats[0].detach().float()
    fl = [(((f.float() - f4) ** 2) * ((f > 0) | (f4 > 0)).float()).sum() for f in feats[1:]]
    total = (1 - alpha) * sum(ce) + alpha * sum(kd) + beta * sum(fl)
    return total, torch.stack(ce + kd + fl)
