"""Segment consensus: the reduction that turns per-segment scores [N, T, K] into one score per clip.

API parity with the reference's ``models/basic_ops.py`` (:9-37): the names ``Identity``,
``SegmentConsensus`` and ``ConsensusModule``, the constructor arguments ``(consensus_type, dim=1)``, the
``'rnn' -> 'identity'`` aliasing, the recorded ``shape`` attribute, and ``None`` for an unknown type.

The training hot path does not go through these modules: the fused classifier head
(``fused.classifier_head`` / ``csrc/head.cu``) takes the mean over T before the GEMV.  They exist for
callers that apply the consensus to their own tensors.
"""
import torch
from torch import nn

# consensus type -> reduction over the segment axis (None = pass through)
_REDUCTIONS = {
    'avg': lambda scores, axis: torch.mean(scores, dim=axis, keepdim=True),
    'identity': None,
}


def _reduce_segments(kind, scores, axis):
    if kind not in _REDUCTIONS:
        return None                       # the reference's behaviour for a type it does not know
    op = _REDUCTIONS[kind]
    return scores if op is None else op(scores, axis)


class Identity(nn.Module):
    """Pass-through (used where the reference swaps a layer out, e.g. ``new_fc`` with dropout 0)."""

    def forward(self, input):
        return input


class SegmentConsensus(nn.Module):
    """One consensus type applied along ``dim``; remembers the last input shape like the reference's Function."""

    def __init__(self, consensus_type, dim=1):
        super().__init__()
        self.consensus_type, self.dim, self.shape = consensus_type, dim, None

    def forward(self, input_tensor):
        self.shape = input_tensor.size()
        return _reduce_segments(self.consensus_type, input_tensor, self.dim)


class ConsensusModule(nn.Module):
    """What ``TSN.consensus`` holds: 'rnn' is treated as 'identity', everything else is forwarded."""

    def __init__(self, consensus_type, dim=1):
        super().__init__()
        self.consensus_type = 'identity' if consensus_type == 'rnn' else consensus_type
        self.dim = dim

    def forward(self, input):
        return _reduce_segments(self.consensus_type, input, self.dim)
