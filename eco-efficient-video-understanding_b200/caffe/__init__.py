"""`caffe` -- drop-in Python package for ECO's hot path on B200.

Mirrors the import surface of caffe_3d/python/caffe/__init__.py:1-7 that ECO's scripts use
(Net, TRAIN/TEST, set_mode_*/set_device); everything runs in libeco_b200.so (sm_100a)."""
from .pycaffe import Net, TRAIN, TEST, set_mode_cpu, set_mode_gpu, set_device, set_logging_disabled, device_count
from ._caffe import Layer, Blob

__all__ = ["Net", "TRAIN", "TEST", "set_mode_cpu", "set_mode_gpu", "set_device", "set_logging_disabled", "Layer",
           "Blob", "device_count"]
from .solver import Solver, SGDSolver, NesterovSolver, get_solver  # noqa: F401
