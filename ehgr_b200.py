"""Alias: ``import ehgr_b200`` == the package directory
``efficient-hand-gesture-recognition-using-multi-task-multi-modal-learning-and-self-distillation_b200``
(whose name is not a Python identifier).  Use attribute access / ``from ehgr_b200 import X``."""
import importlib
import sys

_PKG = "efficient-hand-gesture-recognition-using-multi-task-multi-modal-learning-and-self-distillation_b200"
_pkg = importlib.import_module(_PKG)
sys.modules[__name__] = _pkg
