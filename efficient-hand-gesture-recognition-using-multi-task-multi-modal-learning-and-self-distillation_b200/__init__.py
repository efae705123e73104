"""ehgr_b200 — B200-native (sm_100a) training hot path of the TSM/ACTION-MobileNetV2 gesture
recogniser: hand-written CUDA behind the reference's own operator API.

Import as ``import ehgr_b200`` (alias module at the repo root) — the directory name carries the
reference's full title and is not a Python identifier.
"""
from . import _lib  # noqa: F401
from .temporal_shift import (InplaceShift, TemporalPool, TemporalShift, make_temporal_pool,  # noqa: F401
                             make_temporal_shift, temporal_shift)
from .action import Action  # noqa: F401
from .basic_ops import ConsensusModule, SegmentConsensus  # noqa: F401
from .mobilenet_v2 import InvertedResidual, MobileNetV2, mobilenet_v2  # noqa: F401
from .tsn import TSN  # noqa: F401
from . import losses, train_step, tsn_mtmm  # noqa: F401
from . import action, basic_ops, fused, mobilenet_v2 as mobilenet_v2_module, temporal_shift as temporal_shift_module, tsn  # noqa: F401,E501

__all__ = [
    "TemporalShift", "InplaceShift", "TemporalPool", "make_temporal_shift", "make_temporal_pool",
    "temporal_shift", "Action", "InvertedResidual", "MobileNetV2", "mobilenet_v2", "TSN",
    "ConsensusModule", "SegmentConsensus",
]
