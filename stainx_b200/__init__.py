"""stainx_b200 -- B200-native implementation of StainX's per-pixel stain-normalization hot path.

Public names match ``stainx`` (``src/stainx/__init__.py:L3-7`` of rendeirolab/stainx v0.1.4):
``Reinhard``, ``Macenko``, ``HistogramMatching``, ``StainNormalizerBase``,
``StainNormalizerTransform``.  ``backend="torch_cuda"`` is the only backend; it is served by
``libstainx_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI of
``include/stainx_b200.h``).  There is no CPU path.
"""
from stainx_b200.base import StainNormalizerBase
from stainx_b200.normalizers import HistogramMatching, Macenko, Reinhard
from stainx_b200.transforms import StainNormalizerTransform

__version__ = "0.1.0"

__all__ = ["HistogramMatching", "Macenko", "Reinhard", "StainNormalizerBase", "StainNormalizerTransform", "__version__"]
