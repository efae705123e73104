"""Segment consensus — drop-in for the reference's ``models/basic_ops.py`` (:9-37).

``avg`` is a mean over the segment axis (dim=1, keepdim).  Inside the fused classifier head the mean
over T is folded into the pooled-feature kernel; this module form exists for API parity and for
callers that apply it to their own tensors.
"""
import torch


class Identity(torch.nn.Module):
    def forward(self, input):
        return input


class SegmentConsensus(torch.nn.Module):
    def __init__(self, consensus_type, dim=1):
        super().__init__()
        self.consensus_type = consensus_type
        self.dim = dim
        self.shape = None

    def forward(self, input_tensor):
        self.shape = input_tensor.size()
        if self.consensus_type == 'avg':
            return input_tensor.mean(dim=self.dim, keepdim=True)
        if self.consensus_type == 'identity':
            return input_tensor
        return None


class ConsensusModule(torch.nn.Module):
    def __init__(self, consensus_type, dim=1):
        super().__init__()
        self.consensus_type = consensus_type if consensus_type != 'rnn' else 'identity'
        self.dim = dim

    def forward(self, input):
        return SegmentConsensus(self.consensus_type, self.dim)(input)
