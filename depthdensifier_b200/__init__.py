"""depthdensifier_b200 - B200-native implementation of the DepthDensifier densification hot path.

Keeps the reference package's exports (src/depthdensifier/__init__.py:3-6): ``DepthRefiner``,
``RefinerConfig``, ``__version__``.  Importing the package does not need a GPU; calling any compute
entry point does (there is no CPU fallback)."""

from .depth_refiner import DepthRefiner, RefinerConfig

__version__ = "0.1.0"
__all__ = ["DepthRefiner", "RefinerConfig"]
