"""pyinterpolate_b200 — B200-native (sm_100a) drop-in for pyinterpolate's two data-parallel hot
paths: experimental variogram estimation and ordinary / simple point kriging.

Same call signatures as the reference (src/pyinterpolate/__init__.py:15-24); the NumPy/SciPy
internals are replaced by libpyinterp_b200.so (hand-written CUDA, C ABI in include/pyinterp_b200.h).
There is no CPU fallback: without the CUDA library or a GPU every compute call raises.
"""
from .indicator import ExperimentalIndicatorVariogram, IndicatorKriging, IndicatorVariogramData
from .kriging import (indicator_kriging, interpolate_points, ordinary_and_simple_kriging, ordinary_kriging, simple_kriging,
                      validate_kriging)
from .models import TheoreticalVariogram
from .pipelines import (MultivariateRegression, UniversalKriging, interpolate_points_dask, interpolate_raster,
                        set_dimensions)
from .variogram import (DirectionalVariogram, ExperimentalVariogram, VariogramCloud, build_experimental_variogram,
                        calculate_covariance, calculate_semivariance, point_cloud_semivariance)

__version__ = "0.1.0"
__all__ = ["ExperimentalVariogram", "DirectionalVariogram", "build_experimental_variogram",
           "calculate_semivariance", "calculate_covariance", "point_cloud_semivariance", "VariogramCloud",
           "TheoreticalVariogram",
           "ordinary_kriging", "simple_kriging", "ordinary_and_simple_kriging", "interpolate_points", "indicator_kriging", "validate_kriging",
           "ExperimentalIndicatorVariogram", "IndicatorVariogramData", "IndicatorKriging",
           "interpolate_points_dask", "interpolate_raster", "set_dimensions", "MultivariateRegression", "UniversalKriging"]
