"""OccGridEstimator with the reference's constructor, buffers and method signatures
(perception/nerfacc/nerfacc/estimators/occ_grid.py:13-454), ray marching behind the C-ABI.

Buffers (same names / shapes / dtypes so reference checkpoints' ``occ_grid`` entry loads):
``resolution [3] i32``, ``aabbs [levels, 6] f32``, ``occs [levels * cells] f32``,
``binaries [levels, rx, ry, rz] bool``.
"""
from typing import Callable, List, Optional, Tuple, Union

import torch
from torch import Tensor

from ..grid import _enlarge_aabb, sample_rays, traverse_grids  # noqa: F401
from ..volrend import render_visibility_from_alpha, render_visibility_from_density


class DensityOccEvalFn:
    """The pipeline's ``occ_eval_fn`` (scripts/pipeline.py:376-378: ``query_density(x) * render_step_size``) as an
    object.  It is an ordinary callable, so it can be handed to any estimator; ``OccGridEstimator._update``
    recognises it and runs the whole per-level body -- jittered cell -> density -> EMA-max -- as ONE launch of the
    field kernel (``apnerf_occ_update``) instead of ~10 tensor ops around a density query."""

    def __init__(self, radiance_field, render_step_size: float):
        self.radiance_field = radiance_field
        self.render_step_size = float(render_step_size)

    def __call__(self, x: Tensor) -> Tensor:
        return self.radiance_field.query_density(x) * self.render_step_size


class OccGridEstimator(torch.nn.Module):
    DIM: int = 3

    def __init__(self, roi_aabb: Union[List[int], Tensor], resolution: Union[int, List[int], Tensor] = 128,
                 levels: int = 1, **kwargs) -> None:
        super().__init__()
        if "contraction_type" in kwargs:
            raise ValueError("`contraction_type` is not supported anymore for nerfacc >= 0.4.0.")
        if isinstance(resolution, int):
            resolution = [resolution] * self.DIM
        if isinstance(resolution, (list, tuple)):
            resolution = torch.tensor(resolution, dtype=torch.int32)
        assert isinstance(resolution, Tensor), f"Invalid type: {resolution}!"
        assert resolution.shape[0] == self.DIM, f"Invalid shape: {resolution}!"
        if isinstance(roi_aabb, (list, tuple)):
            roi_aabb = torch.tensor(roi_aabb, dtype=torch.float32)
        assert isinstance(roi_aabb, Tensor), f"Invalid type: {roi_aabb}!"
        assert roi_aabb.shape[0] == self.DIM * 2, f"Invalid shape: {roi_aabb}!"

        aabbs = torch.stack([_enlarge_aabb(roi_aabb, 2 ** i) for i in range(levels)], dim=0)
        self.cells_per_lvl = int(resolution.prod().item())
        self.levels = levels
        self.register_buffer("resolution", resolution)
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(self.levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros([levels] + resolution.tolist(), dtype=torch.bool))
        res = resolution.tolist()
        coords = torch.stack(torch.meshgrid([torch.arange(r, dtype=torch.long) for r in res], indexing="ij"), dim=-1)
        self.register_buffer("grid_coords", coords.reshape(self.cells_per_lvl, self.DIM), persistent=False)
        self.register_buffer("grid_indices", torch.arange(self.cells_per_lvl), persistent=False)

    @property
    def device(self) -> torch.device:
        return self.occs.device

    @torch.no_grad()
    def sampling(
        self, rays_o: Tensor, rays_d: Tensor, sigma_fn: Optional[Callable] = None,
        alpha_fn: Optional[Callable] = None, near_plane: float = 0.0, far_plane: float = 1e10,
        t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None, render_step_size: float = 1e-3,
        early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, stratified: bool = False,
        cone_angle: float = 0.0, depth: Optional[Tensor] = None,
    ) -> Tuple[Tensor, Tensor, Tensor]:
        """occ_grid.py:80-238.  ``depth`` is accepted and ignored, as in the reference fork."""
        near_planes = torch.full_like(rays_o[..., 0], fill_value=near_plane)
        far_planes = torch.full_like(rays_o[..., 0], fill_value=far_plane)
        if t_min is not None:
            near_planes = torch.clamp(near_planes, min=t_min)
        if t_max is not None:
            far_planes = torch.clamp(far_planes, max=t_max)
        if stratified:
            near_planes += torch.rand_like(near_planes) * render_step_size
        # traverse_grids + the two boolean compactions of the reference (:117-131) as one count / scan / fill
        ray_indices, t_starts, t_ends, packed_info = sample_rays(
            rays_o, rays_d, self.binaries, self.aabbs, near_planes, far_planes, render_step_size, cone_angle)

        if (alpha_thre > 0.0 or early_stop_eps > 0.0) and (sigma_fn is not None or alpha_fn is not None):
            alpha_thre = min(alpha_thre, self._occs_mean())
            if sigma_fn is not None:
                if t_starts.shape[0] != 0:
                    sigmas = sigma_fn(t_starts, t_ends, ray_indices)
                else:
                    sigmas = torch.empty((0,), device=t_starts.device)
                assert sigmas.shape == t_starts.shape, "sigmas must have shape of (N,)! Got {}".format(sigmas.shape)
                masks = render_visibility_from_density(
                    t_starts=t_starts, t_ends=t_ends, sigmas=sigmas, packed_info=packed_info,
                    early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
            else:
                if t_starts.shape[0] != 0:
                    alphas = alpha_fn(t_starts, t_ends, ray_indices)
                else:
                    alphas = torch.empty((0,), device=t_starts.device)
                assert alphas.shape == t_starts.shape, "alphas must have shape of (N,)! Got {}".format(alphas.shape)
                masks = render_visibility_from_alpha(
                    alphas=alphas, packed_info=packed_info, early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
            keep = masks.nonzero(as_tuple=True)[0]  # ONE host sync for the three compactions (the reference has three)
            ray_indices, t_starts, t_ends = ray_indices[keep], t_starts[keep], t_ends[keep]
        return ray_indices, t_starts, t_ends

    def _occs_mean(self) -> float:
        """``self.occs.mean().item()`` (occ_grid.py:199) without a host synchronisation per call: the occupancy values
        only change when the grid is updated (every 16 training steps), so the mean is read back once per version of
        the buffer."""
        key = (self.occs.data_ptr(), self.occs._version)
        if getattr(self, "_occs_mean_key", None) != key:
            self._occs_mean_val = self.occs.mean().item()
            self._occs_mean_key = key
        return self._occs_mean_val

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2,
                             ema_decay: float = 0.95, warmup_steps: int = 256, n: int = 16) -> None:
        """occ_grid.py:240-276."""
        if not self.training:
            raise RuntimeError(
                "You should only call this function only during training. "
                "Please call _update() directly if you want to update the field during inference.")
        if step % n == 0 and self.training:
            self._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre, ema_decay=ema_decay,
                         warmup_steps=warmup_steps)

    @torch.no_grad()
    def _get_all_cells(self) -> List[Tensor]:
        out = []
        for lvl in range(self.levels):
            cell_ids = lvl * self.cells_per_lvl + self.grid_indices
            out.append(self.grid_indices[self.occs[cell_ids] >= 0.0])
        return out

    @torch.no_grad()
    def _sample_uniform_and_occupied_cells(self, n: int) -> List[Tensor]:
        out = []
        for lvl in range(self.levels):
            uniform = torch.randint(self.cells_per_lvl, (n,), device=self.device)
            uniform = uniform[self.occs[lvl * self.cells_per_lvl + uniform] >= 0.0]
            occupied = torch.nonzero(self.binaries[lvl].flatten())[:, 0]
            if n < len(occupied):
                occupied = occupied[torch.randint(len(occupied), (n,), device=self.device)]
            out.append(torch.cat([uniform, occupied], dim=0))
        return out

    @torch.no_grad()
    def _update_level_fused(self, lvl: int, indices: Tensor, jitter: Tensor, backup: Tensor,
                            occ_eval_fn: DensityOccEvalFn, ema_decay: float) -> None:
        """occ_grid.py:396-434 for one level in one kernel (csrc/field_kernel.cuh, occupancy-update mode)."""
        import ctypes

        import numpy as np

        from ..._lib import call

        f = occ_eval_fn.radiance_field
        weights, table = f._packed()
        aabb_host = f.aabb_host()
        lvl_aabb = np.ascontiguousarray(self.aabbs[lvl].detach().cpu().numpy(), dtype=np.float32)
        res = [int(v) for v in self.resolution.tolist()]
        lo, hi = lvl * self.cells_per_lvl, (lvl + 1) * self.cells_per_lvl
        with torch.cuda.device(self.occs.device):
            call("apnerf_occ_update", int(indices.numel()), indices.contiguous(), jitter.contiguous(),
                 lvl_aabb.ctypes.data_as(ctypes.c_void_p), res[0], res[1], res[2], backup[lo:hi], self.occs[lo:hi],
                 float(occ_eval_fn.render_step_size), float(ema_decay), aabb_host.ctypes.data_as(ctypes.c_void_p),
                 f.n_levels, f._meta.ctypes.data_as(ctypes.c_void_p), table, weights)

    @torch.no_grad()
    def _update(self, step: int, occ_eval_fn: Callable, occ_thre: float = 0.01, ema_decay: float = 0.95,
                warmup_steps: int = 256) -> None:
        """EMA-max occupancy update + binarisation, occ_grid.py:377-437 (incl. the fork's NaN
        restore :405, :430-434)."""
        if step < warmup_steps:
            lvl_indices = self._get_all_cells()
        else:
            lvl_indices = self._sample_uniform_and_occupied_cells(self.cells_per_lvl // 4)
        fused = (isinstance(occ_eval_fn, DensityOccEvalFn) and hasattr(occ_eval_fn.radiance_field, "_packed")
                 and self.occs.is_cuda)
        for lvl, indices in enumerate(lvl_indices):
            grid_coords = self.grid_coords[indices]
            jitter = torch.rand_like(grid_coords, dtype=torch.float32)
            backup = torch.clone(self.occs)
            if fused:
                self._update_level_fused(lvl, indices, jitter, backup, occ_eval_fn, ema_decay)
                continue
            x = (grid_coords + jitter) / self.resolution
            x = self.aabbs[lvl, :3] + x * (self.aabbs[lvl, 3:] - self.aabbs[lvl, :3])
            occ = occ_eval_fn(x).squeeze(-1)
            cell_ids = lvl * self.cells_per_lvl + indices
            self.occs[cell_ids] = torch.maximum(self.occs[cell_ids] * ema_decay, occ)
            bad = torch.isnan(self.occs)
            if bad.any():
                self.occs[bad] = backup[bad]
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)
