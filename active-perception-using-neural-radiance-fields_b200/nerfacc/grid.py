"""ray_aabb_intersect / traverse_grids with the reference's Python signatures
(perception/nerfacc/nerfacc/grid.py:13-51, 93-192); the host logic of the reference's C++
launcher (csrc/grid.cu:320-474, csrc/include/data_spec.hpp:86-106) lives here in PyTorch,
the kernels are behind the C-ABI."""
from typing import Optional, Tuple

import torch
from torch import Tensor

from .._lib import LIB, call, require_cuda
from .data_specs import RayIntervals, RaySamples


@torch.no_grad()
def ray_aabb_intersect(rays_o: Tensor, rays_d: Tensor, aabbs: Tensor, near_plane: float = -float("inf"),
                       far_plane: float = float("inf"), miss_value: float = float("inf")
                       ) -> Tuple[Tensor, Tensor, Tensor]:
    assert rays_o.ndim == 2 and rays_o.shape[-1] == 3
    assert rays_d.ndim == 2 and rays_d.shape[-1] == 3
    assert aabbs.ndim == 2 and aabbs.shape[-1] == 6
    require_cuda(rays_o, rays_d, aabbs)
    rays_o, rays_d, aabbs = rays_o.float().contiguous(), rays_d.float().contiguous(), aabbs.float().contiguous()
    n, m = rays_o.shape[0], aabbs.shape[0]
    t_mins = torch.empty((n, m), device=rays_o.device, dtype=torch.float32)
    t_maxs = torch.empty((n, m), device=rays_o.device, dtype=torch.float32)
    hits = torch.empty((n, m), device=rays_o.device, dtype=torch.bool)
    with torch.cuda.device(rays_o.device):
        call("apnerf_ray_aabb_intersect", n, rays_o, rays_d, m, aabbs, float(near_plane), float(far_plane),
             float(miss_value), t_mins, t_maxs, hits)
    return t_mins, t_maxs, hits


def _exclusive_scan_i64(cnts: Tensor) -> Tuple[Tensor, Tensor]:
    """chunk_cnts -> (chunk_starts, total[1]) on the device."""
    n = cnts.numel()
    starts = torch.empty_like(cnts)
    total = torch.zeros(1, device=cnts.device, dtype=torch.int64)
    scratch = torch.empty(int(LIB.raw("apnerf_scan_scratch_elems")(n)), device=cnts.device, dtype=torch.int64)
    call("apnerf_exclusive_scan_i64", n, cnts, starts, total, scratch)
    return starts, total


@torch.no_grad()
def traverse_grids(
    rays_o: Tensor, rays_d: Tensor, binaries: Tensor, aabbs: Tensor,
    near_planes: Optional[Tensor] = None, far_planes: Optional[Tensor] = None,
    step_size: Optional[float] = 1e-3, cone_angle: Optional[float] = 0.0,
    traverse_steps_limit: Optional[int] = None, over_allocate: Optional[bool] = False,
    rays_mask: Optional[Tensor] = None, t_sorted: Optional[Tensor] = None,
    t_indices: Optional[Tensor] = None, hits: Optional[Tensor] = None,
) -> Tuple[RayIntervals, RaySamples, Tensor]:
    require_cuda(rays_o, rays_d, binaries, aabbs)
    if near_planes is None:
        near_planes = torch.zeros_like(rays_o[:, 0])
    if far_planes is None:
        far_planes = torch.full_like(rays_o[:, 0], float("inf"))
    if rays_mask is None:
        rays_mask = torch.ones_like(rays_o[:, 0], dtype=torch.bool)
    if traverse_steps_limit is None:
        traverse_steps_limit = -1
    if over_allocate:
        assert traverse_steps_limit > 0, "traverse_steps_limit must be set if over_allocate is True."
    if t_sorted is None or t_indices is None or hits is None:
        t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, aabbs)
        t_sorted, t_indices = torch.sort(torch.cat([t_mins, t_maxs], dim=-1), dim=-1)

    # the kernel reads raw float32 / int64 / 1-byte pointers: normalise dtypes here
    rays_o, rays_d = rays_o.float().contiguous(), rays_d.float().contiguous()
    rays_mask, binaries, aabbs = rays_mask.bool().contiguous(), binaries.bool().contiguous(), aabbs.float().contiguous()
    t_sorted, t_indices, hits = t_sorted.float().contiguous(), t_indices.to(torch.int64).contiguous(), hits.bool().contiguous()
    near_planes, far_planes = near_planes.float().contiguous(), far_planes.float().contiguous()
    dev = rays_o.device
    n_rays, n_grids = rays_o.shape[0], binaries.shape[0]
    rx, ry, rz = (int(s) for s in binaries.shape[1:])
    i64 = dict(device=dev, dtype=torch.int64)
    f32 = dict(device=dev, dtype=torch.float32)
    b8 = dict(device=dev, dtype=torch.bool)
    terminate_planes = torch.empty(n_rays, **f32)

    def launch(mask, first_pass, iv, sm, term):
        call("apnerf_traverse_grids", n_rays, rays_o, rays_d, mask, n_grids, rx, ry, rz, binaries, aabbs, hits,
             t_sorted, t_indices, near_planes, far_planes, float(step_size), float(cone_angle),
             int(traverse_steps_limit), 1 if first_pass else 0,
             iv.get("vals"), iv.get("ray_indices"), iv.get("is_left"), iv.get("is_right"),
             iv.get("chunk_starts"), iv["chunk_cnts"],
             sm.get("vals"), sm.get("ray_indices"), sm.get("is_valid"), sm.get("chunk_starts"), sm["chunk_cnts"],
             term)

    def alloc(spec, masks, valid):  # RaySegmentsSpec::memalloc_data_from_chunk (zero-initialised)
        starts, total = _exclusive_scan_i64(spec["chunk_cnts"])
        n_edges = int(total.item())  # the one host sync the reference also has (data_spec.hpp:91)
        spec["chunk_starts"] = starts
        spec["vals"] = torch.zeros(n_edges, **f32)
        spec["ray_indices"] = torch.zeros(n_edges, **i64)
        if masks:
            spec["is_left"] = torch.zeros(n_edges, **b8)
            spec["is_right"] = torch.zeros(n_edges, **b8)
        if valid:
            spec["is_valid"] = torch.zeros(n_edges, **b8)

    with torch.cuda.device(dev):
        iv, sm = {}, {}
        if over_allocate:  # csrc/grid.cu:364-404
            iv["chunk_cnts"] = torch.full((n_rays,), traverse_steps_limit * 2, **i64) * rays_mask
            alloc(iv, True, False)
            sm["chunk_cnts"] = torch.full((n_rays,), traverse_steps_limit, **i64) * rays_mask
            alloc(sm, False, True)
            launch(rays_mask, False, iv, sm, terminate_planes)
            iv["chunk_starts"], _ = _exclusive_scan_i64(iv["chunk_cnts"])
            sm["chunk_starts"], _ = _exclusive_scan_i64(sm["chunk_cnts"])
        else:  # csrc/grid.cu:405-470 : count pass, allocate, fill pass (rays_mask ignored)
            iv["chunk_cnts"] = torch.empty(n_rays, **i64)
            sm["chunk_cnts"] = torch.empty(n_rays, **i64)
            launch(None, True, iv, sm, None)
            alloc(iv, True, False)
            alloc(sm, False, True)
            launch(None, False, iv, sm, terminate_planes)

    intervals = RayIntervals(
        vals=iv["vals"], packed_info=torch.stack([iv["chunk_starts"], iv["chunk_cnts"]], -1),
        ray_indices=iv["ray_indices"], is_left=iv["is_left"], is_right=iv["is_right"])
    samples = RaySamples(
        vals=sm["vals"], packed_info=torch.stack([sm["chunk_starts"], sm["chunk_cnts"]], -1),
        ray_indices=sm["ray_indices"], is_valid=sm["is_valid"])
    return intervals, samples, terminate_planes


@torch.no_grad()
def sample_rays(rays_o: Tensor, rays_d: Tensor, binaries: Tensor, aabbs: Tensor, near_planes: Tensor, far_planes: Tensor,
                step_size: float, cone_angle: float):
    """What ``OccGridEstimator.sampling`` takes from ``traverse_grids`` (occ_grid.py:117-131) -- the packed samples
    ``(ray_indices i64, t_starts, t_ends, packed_info [n_rays, 2] i64)`` -- through ``apnerf_sample_rays``: count,
    scan, fill; ONE host synchronisation (the sample total) where the interval route has four."""
    require_cuda(rays_o, rays_d, binaries, aabbs)
    rays_o, rays_d = rays_o.float().contiguous(), rays_d.float().contiguous()
    binaries, aabbs = binaries.bool().contiguous(), aabbs.float().contiguous()
    near_planes, far_planes = near_planes.float().contiguous(), far_planes.float().contiguous()
    n_rays, n_grids = rays_o.shape[0], binaries.shape[0]
    rx, ry, rz = (int(s) for s in binaries.shape[1:])
    dev = rays_o.device
    t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, aabbs)
    if n_grids > 1:
        t_sorted, t_indices = torch.sort(torch.cat([t_mins, t_maxs], dim=-1), dim=-1)
        t_sorted, t_indices = t_sorted.contiguous(), t_indices.contiguous()
    else:
        t_sorted, t_indices = torch.cat([t_mins, t_maxs], dim=-1).contiguous(), None
    cnts = torch.empty(n_rays, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        args = (n_rays, rays_o, rays_d, n_grids, rx, ry, rz, binaries, aabbs, hits.contiguous(), t_sorted, t_indices,
                near_planes, far_planes, float(step_size), float(cone_angle), -1)
        call("apnerf_sample_rays", *args, None, cnts, None, None, None)
        starts, total = _exclusive_scan_i64(cnts)
        n = int(total.item())
        ray_indices = torch.empty(n, device=dev, dtype=torch.int64)
        t_starts, t_ends = torch.empty(n, device=dev), torch.empty(n, device=dev)
        if n:
            call("apnerf_sample_rays", *args, starts, cnts, ray_indices, t_starts, t_ends)
    return ray_indices, t_starts, t_ends, torch.stack([starts, cnts], -1)


def _enlarge_aabb(aabb, factor: float) -> Tensor:
    center = (aabb[:3] + aabb[3:]) / 2
    extent = (aabb[3:] - aabb[:3]) / 2
    return torch.cat([center - extent * factor, center + extent * factor])
