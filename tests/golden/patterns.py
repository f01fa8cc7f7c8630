"""Deterministic stand-in scene functions shared by the golden generator (CPU, reference Python) and the
GPU tests (CUDA mirrors).  Every value is built from operations that are exact in fp32 on both sides
(power-of-two scalings, floor, integer remainders), so the comparisons can be bit-exact."""
import torch


def sigma_pattern(t_starts, t_ends, ray_indices):
    """Density that depends on the ray and on the 1/64-quantised distance: 0, 4 or 8."""
    return ((ray_indices + (t_starts * 64.0).floor().long()) % 3).float() * 4.0


def rgb_sigma_pattern(t_starts, t_ends, ray_indices):
    r = (ray_indices % 8).float() / 8.0
    g = ((t_starts * 32.0).floor() % 4.0) / 4.0
    b = torch.full_like(r, 0.5)
    return torch.stack([r, g, b], -1), sigma_pattern(t_starts, t_ends, ray_indices)


def occ_pattern(aabb, resolution):
    """occ_eval_fn for OccGridEstimator._update: a function of the cell the point falls in (points are cell
    centres when torch.rand_like is patched to 0.5), multiples of 1/64 in [0, 4/64]."""
    lo, ext = aabb[:3], aabb[3:] - aabb[:3]

    def fn(x):
        cell = ((x - lo) / ext * resolution).floor().long()
        return ((cell[:, 0] + 2 * cell[:, 1] + 3 * cell[:, 2]) % 5).float()[:, None] / 64.0

    return fn


def half_like(t, dtype=None):
    """Replacement for torch.rand_like inside _update / stratified sampling: the constant 0.5."""
    return torch.full(t.shape, 0.5, dtype=dtype or t.dtype, device=t.device)
