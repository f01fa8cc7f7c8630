from .occ_grid import DensityOccEvalFn, OccGridEstimator

__all__ = ["OccGridEstimator", "DensityOccEvalFn"]
