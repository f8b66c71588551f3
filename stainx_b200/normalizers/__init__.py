from stainx_b200.normalizers.histogram_matching import HistogramMatching
from stainx_b200.normalizers.macenko import Macenko
from stainx_b200.normalizers.reinhard import Reinhard

__all__ = ["HistogramMatching", "Macenko", "Reinhard"]
