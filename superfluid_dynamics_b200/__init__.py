"""superfluid_dynamics_b200 -- B200-native Roberts (1983) boundary-integral RK4 step behind CuSuperHelium's interface.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/roberts_b200.h) and api.py (host-side mirror of
the reference's solver / stepper classes).  Importing the package does not import torch; ``superfluid_dynamics_b200.api``
does (device memory, streams).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
