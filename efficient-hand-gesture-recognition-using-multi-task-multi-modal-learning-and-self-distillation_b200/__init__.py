"""ehgr_b200 — B200-native (sm_100a) training hot path of the TSM/ACTION-MobileNetV2 gesture
recogniser: hand-written CUDA behind the reference's own operator API.

Import as ``import ehgr_b200`` (alias module at the repo root) — the directory name carries the
reference's full title and is not a Python identifier.  Sub-modules are exposed as attributes with a
``_module`` suffix where a re-exported function has the same name (``temporal_shift``).
"""
import importlib as _importlib

from . import _lib  # noqa: F401

temporal_shift_module = _importlib.import_module(__name__ + ".temporal_shift")
mobilenet_v2_module = _importlib.import_module(__name__ + ".mobilenet_v2")
action = _importlib.import_module(__name__ + ".action")
basic_ops = _importlib.import_module(__name__ + ".basic_ops")
fused = _importlib.import_module(__name__ + ".fused")
resnet_ops = _importlib.import_module(__name__ + ".resnet_ops")
tsn = _importlib.import_module(__name__ + ".tsn")
losses = _importlib.import_module(__name__ + ".losses")
train_step = _importlib.import_module(__name__ + ".train_step")
tsn_mtmm = _importlib.import_module(__name__ + ".tsn_mtmm")
tsn_sd = _importlib.import_module(__name__ + ".tsn_sd")
tsn_mtmm_sd = _importlib.import_module(__name__ + ".tsn_mtmm_sd")

from .temporal_shift import (InplaceShift, TemporalPool, TemporalShift, make_temporal_pool,  # noqa: E402,F401
                             make_temporal_shift, temporal_shift)
from .action import Action  # noqa: E402,F401
from .basic_ops import ConsensusModule, SegmentConsensus  # noqa: E402,F401
from .mobilenet_v2 import InvertedResidual, MobileNetV2, mobilenet_v2  # noqa: E402,F401
from .tsn import TSN  # noqa: E402,F401

__all__ = [
    "TemporalShift", "InplaceShift", "TemporalPool", "make_temporal_shift", "make_temporal_pool",
    "temporal_shift", "Action", "InvertedResidual", "MobileNetV2", "mobilenet_v2", "TSN",
    "ConsensusModule", "SegmentConsensus",
]
