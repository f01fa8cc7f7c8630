"""Drop-in mirror of the reference's vendored ``nerfacc`` package surface that the render /
score / train hot path uses (perception/nerfacc/nerfacc/__init__.py), backed by libapnerf.so."""
from .data_specs import RayIntervals, RaySamples
from .estimators.occ_grid import DensityOccEvalFn, OccGridEstimator
from .grid import ray_aabb_intersect, traverse_grids
from .pack import pack_info
from .scan import exclusive_sum, inclusive_sum
from .volrend import (
    accumulate_along_rays,
    accumulate_along_rays_,
    render_transmittance_from_density,
    render_visibility_from_density,
    render_weight_from_density,
    rendering,
)

__all__ = [
    "RayIntervals", "RaySamples", "OccGridEstimator", "DensityOccEvalFn", "ray_aabb_intersect", "traverse_grids", "pack_info",
    "exclusive_sum", "inclusive_sum", "accumulate_along_rays", "accumulate_along_rays_",
    "render_transmittance_from_density", "render_visibility_from_density", "render_weight_from_density",
    "rendering",
]
