"""Backend classes (only ``torch_cuda`` exists in this build)."""
from stainx_b200.backends.torch_cuda_backend import CUDA_AVAILABLE, HistogramMatchingCUDA, MacenkoCUDA, ReinhardCUDA

__all__ = ["CUDA_AVAILABLE", "HistogramMatchingCUDA", "MacenkoCUDA", "ReinhardCUDA"]
