"""B200-native (sm_100a) render + score hot path for active perception with NeRFs.

A drop-in for the reference's nerfacc ops, NGP radiance field, test-mode renderers and
predictive-information scorer; all device work is hand-written CUDA behind the C-ABI library
``libapnerf.so`` (include/apnerf.h).  There is no CPU or PyTorch fallback.
"""
from . import _lib, data_proc, nerfacc, pipeline, radiance_fields, render, scoring, synthetic, training  # noqa: F401
from .pipeline import ActiveNeRFMapper  # noqa: F401
from .data_proc import Dataset  # noqa: F401
from .nerfacc import OccGridEstimator  # noqa: F401
from .radiance_fields import NGPRadianceField  # noqa: F401

from .render import (  # noqa: F401
    FusedRenderer,
    Rays,
    render_image_with_occgrid,
    render_image_with_occgrid_test,
    render_image_with_occgrid_with_depth_guide,
    render_probablistic_image_with_occgrid_test,
    sem_rendering,
)
from .scoring import PredictiveInformationScorer, probablistic_uncertainty, trajector_uncertainty  # noqa: F401

__all__ = ["nerfacc", "radiance_fields", "render", "scoring", "synthetic", "OccGridEstimator", "NGPRadianceField",
           "FusedRenderer", "Rays", "render_image_with_occgrid", "render_image_with_occgrid_test", "render_probablistic_image_with_occgrid_test",
           "render_image_with_occgrid_with_depth_guide", "sem_rendering", "training", "PredictiveInformationScorer", "probablistic_uncertainty", "trajector_uncertainty", "Dataset", "data_proc", "_lib", "pipeline", "ActiveNeRFMapper"]
