"""barefoot_framework_b200: B200-native (sm_100a) implementation of BAREFOOT's
data-parallel inner loop behind the reference's Python API.

Host modules mirror the reference's names (gpModel, reificationFusion,
acquisitionFunc, util, barefoot); every numeric call goes through the C ABI in
include/barefoot_sm100.h (hand-written CUDA, no CPU fallback)."""
from . import _lib  # noqa: F401

__all__ = ["_lib", "engine", "dist", "gpModel", "reificationFusion", "acquisitionFunc", "util", "barefoot", "subProcess"]
