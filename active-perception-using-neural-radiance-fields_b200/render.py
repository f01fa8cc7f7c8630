"""Render glue with the reference's call surface (perception/models/utils.py).

Two implementations of the test-mode renderers live here:

* ``render_image_with_occgrid_test`` / ``render_probablistic_image_with_occgrid_test`` -- the
  drop-ins.  For a single-level occupancy grid (the pipeline's configuration) they run the
  device-driven renderer (csrc/render.cu): the reference's per-call marching schedule is
  replayed on the GPU with no host synchronisation inside the loop: per iteration ``schedule ->
  march -> field (hash grid + MLPs) -> composite``.  ``FusedRenderer.render(fuse_compositor=True)``
  selects the variant that runs the compositor INSIDE the field kernel's epilogue (tile-aware march,
  ordered live-list compaction; per-sample network outputs never leave the SM).  Both are tested
  against each other and the oracle; the stand-alone compositor is the default because it measured
  ~7 % faster in round 1 (profiles/r01_fused_compositor.md).
* ``*_unfused`` -- the same algorithm written op by op against the drop-in nerfacc ops, with
  the reference's host-side control flow (``.item()`` per iteration).  Used for multi-level
  grids and as the on-GPU cross-check of the fused path.

Train-mode glue (``render_image_with_occgrid_with_depth_guide``, ``sem_rendering``) follows
utils.py:63-219, 362-461 on top of ``OccGridEstimator.sampling`` and the packed volrend ops.
"""
import collections
import time
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from ._lib import PreparedCall, call, require_cuda
from .nerfacc import (
    OccGridEstimator,
    accumulate_along_rays,
    accumulate_along_rays_,
    ray_aabb_intersect,
    render_weight_from_density,
    traverse_grids,
)

Rays = collections.namedtuple("Rays", ("origins", "viewdirs"))  # perception/models/datasets/utils.py:7-12

ST_RGB, ST_OPA, ST_DEPTH, ST_RGBVAR, ST_DVAR, ST_SEM = 0, 3, 4, 5, 8, 9


def namedtuple_map(fn, tup):
    return type(tup)(*(None if x is None else fn(x) for x in tup))


# ------------------------------------------------------------------------------------------
# Fused, device-driven renderer
# ------------------------------------------------------------------------------------------
class FusedRenderer:
    """Owns the HBM working set of one batch of calls (views x one ensemble member) and
    enqueues the per-iteration kernel sequence schedule -> march -> field -> composite."""
    host_blocked_s = 0.0  # seconds the enqueuing thread spent waiting for the device (throttle), all renderers: measurement

    def __init__(self, device, n_sem: int):
        self.device = torch.device(device)
        self.n_sem = int(n_sem)
        self.n_state = 9 + self.n_sem
        self._cap_rays = 0
        self._cap_samples = 0
        self._cap_calls = 0
        self._pinned = None

    def _ensure(self, n_rays, n_calls, s_cap, need_rows=False):
        dev = self.device
        if n_rays > self._cap_rays:
            self.t_min = torch.empty(n_rays, device=dev)
            self.t_max = torch.empty(n_rays, device=dev)
            self.hit = torch.empty(n_rays, device=dev, dtype=torch.uint8)
            self.near = torch.empty(n_rays, device=dev)
            self.alive = [torch.empty(n_rays, device=dev, dtype=torch.int32) for _ in range(2)]
            self.entry_base = torch.empty(n_rays, device=dev, dtype=torch.int32)
            self.entry_cnt = torch.empty(n_rays, device=dev, dtype=torch.int32)
            self._cap_rays = n_rays
        if s_cap > self._cap_samples:
            self.s_ray = torch.empty(s_cap, device=dev, dtype=torch.int32)
            self.s_ts = torch.empty(s_cap, device=dev)
            self.s_te = torch.empty(s_cap, device=dev)
            self.s_x = torch.empty((s_cap, 4), device=dev)  # aabb-normalised sample points (marcher -> field kernel)
            self._cap_samples = s_cap
        if need_rows and s_cap > getattr(self, "_cap_rows", 0):
            self.rows = torch.empty((s_cap, 40), device=dev, dtype=torch.float16)  # raw fp16 network outputs
            self._cap_rows = s_cap
        if n_calls > self._cap_calls:
            self.n_alive_acc = torch.empty(n_calls, device=dev, dtype=torch.int32)
            self.n_samp = torch.empty(n_calls, device=dev, dtype=torch.int32)
            self.iter_samples = torch.empty(n_calls, device=dev, dtype=torch.int32)
            self.total_samples = torch.empty(n_calls, device=dev, dtype=torch.int32)
            self._cap_calls = n_calls
        if n_rays > getattr(self, "_cap_flags", 0):
            self.keep_flag = torch.zeros(n_rays, device=dev, dtype=torch.uint8)
            self.chain = torch.zeros(n_rays // 2048 + 2, device=dev, dtype=torch.int64)
            self._cap_flags = n_rays
        if s_cap > getattr(self, "_cap_cnt", 0):
            self.s_cnt = torch.zeros(s_cap, device=dev, dtype=torch.uint8)
            self._cap_cnt = s_cap
        if not hasattr(self, "counters"):
            self.counters = torch.zeros(16, device=dev, dtype=torch.int32)
            self._tag = 0

    def reserve(self, n_rays: int, n_calls: int, cone_angle: float = 0.004) -> None:
        """Allocate the working set for renders of up to n_rays rays now (stand-alone compositor layout), so that no
        render grows it later: a growth in the middle of a step is a cudaMalloc, i.e. a device-wide stall."""
        self._ensure(int(n_rays), int(n_calls), int(n_rays) * (1 if cone_angle == 0 else 4), need_rows=True)

    @torch.no_grad()
    def render(self, *args, **kwargs) -> Tensor:
        """Run ``render_iter`` to completion and return the state (see there)."""
        state = None
        for state in self.render_iter(*args, **kwargs):
            pass
        return state

    @torch.no_grad()
    def render_iter(self, radiance_field, estimator: OccGridEstimator, rays_o: Tensor, rays_d: Tensor,
               rays_per_call: int, *, max_samples: int = 1024, near_plane: float = 0.0, far_plane: float = 1e10,
               render_step_size: float = 1e-3, cone_angle: float = 0.0, alpha_thre: float = 0.0,
               early_stop_eps: float = 1e-4, probabilistic: bool = True, state: Optional[Tensor] = None,
               poll_every: int = 4, debug_hook: Optional[Callable] = None, fuse_compositor: bool = False,
               ray_counts: Optional[Tensor] = None, call_rows: Optional[Tensor] = None,
               min_samples: Optional[int] = None):
        """Render n_rays = n_calls * rays_per_call rays into the state [9 + C, n_rays] (un-finalised: see
        ``finalize``).  A generator: it enqueues one marching iteration on the CURRENT stream per ``next()``
        and yields the state tensor, so a caller can interleave several renders on different streams.
        The device decides everything; the host only (a) stays at most ~2 * poll_every iterations ahead of
        the GPU and (b) looks at the live-ray counter (pinned memory, copies enqueued every ``poll_every``
        iterations) to stop enqueuing once every ray has terminated.  ``ray_counts`` (int32 [2, n_rays], zeroed
        by the caller) accumulates per ray the samples evaluated / composited: test instrumentation.  ``call_rows``
        (int32 [n_calls], zeroed by the caller) accumulates per call the sample rows sent through the field (the
        scheduler's cost probe).  ``min_samples`` overrides the reference's lower bound of the per-iteration sample
        count (1 without a cone angle, else 4; utils.py:894): only the cost probe does that, to finish in fewer,
        larger iterations -- the composited result is then NOT the reference's."""
        import ctypes

        ahead = 2
        require_cuda(rays_o, rays_d, estimator.binaries)
        assert estimator.binaries.shape[0] == 1, "the fused renderer handles single-level occupancy grids"
        assert radiance_field.num_semantic_classes == self.n_sem
        rays_o = rays_o.reshape(-1, 3).contiguous().float()
        rays_d = rays_d.reshape(-1, 3).contiguous().float()
        n_rays = rays_o.shape[0]
        assert n_rays % rays_per_call == 0
        n_calls = n_rays // rays_per_call
        if min_samples is None:
            min_samples = 1 if cone_angle == 0 else 4
        min_samples = int(min_samples)
        # rows of the per-iteration sample list: <= min_samples per ray (utils.py:902); the fused layout
        # pads every warp's run to whole 128-row tiles (<= 2x for the worst non-power-of-two n)
        # + up to 127 padding rows per 32-ray warp; the kernel refuses (flag) rather than overflow
        s_cap = n_rays * (min_samples * 2 + 4) + 4096 if fuse_compositor else n_rays * min_samples
        self._ensure(n_rays, n_calls, s_cap, need_rows=not fuse_compositor)
        if state is None:
            state = torch.empty((self.n_state, n_rays), device=self.device)
        binaries = estimator.binaries.contiguous()
        aabbs = estimator.aabbs.contiguous().float()
        rx, ry, rz = (int(v) for v in binaries.shape[1:])
        n_words = (rx * ry * rz + 31) // 32  # the marcher's one-bit-per-cell copy of the grid, rebuilt by render_init
        if getattr(self, "occ_bits", None) is None or self.occ_bits.numel() < n_words:
            self.occ_bits = torch.empty(n_words, device=self.device, dtype=torch.int32)
        weights, table = radiance_field._packed()
        aabb_host = radiance_field.aabb_host()
        meta = radiance_field._meta
        opc_thre = float(np.float32(1 - early_stop_eps))
        max_iters = (max_samples + min_samples - 1) // min_samples
        if self._pinned is None or self._pinned.numel() < max_iters:
            self._pinned = torch.empty(max_iters, dtype=torch.int32, pin_memory=True)
        events = []
        with torch.cuda.device(self.device):
            call("apnerf_render_init", n_rays, rays_per_call, rays_o, rays_d, rx, ry, rz, binaries, aabbs,
                 float(near_plane), self.n_state, state, self.t_min, self.t_max, self.hit, self.near, self.alive[1],
                 self.n_alive_acc, self.iter_samples, self.total_samples, n_calls, self.counters, self.occ_bits)
            # the per-iteration launch sequence, marshalled once (two variants: the live lists ping-pong)
            aabb_p = aabb_host.ctypes.data_as(ctypes.c_void_p)
            meta_p = meta.ctypes.data_as(ctypes.c_void_p)
            seq = []
            for parity in range(2):
                cur, nxt = self.alive[(parity + 1) % 2], self.alive[parity % 2]
                steps = [PreparedCall("apnerf_render_schedule", n_calls, rays_per_call, int(max_samples), min_samples,
                                      self.n_alive_acc, self.n_samp, self.iter_samples, self.counters, call_rows)]
                if fuse_compositor:
                    # three launches: tile-aware march -> field + compositor fused -> ordered compaction
                    steps.append(PreparedCall(
                        "apnerf_render_march_tiles", n_rays, rays_per_call, cur, self.n_samp, rays_o, rays_d, rx, ry,
                        rz, binaries, aabbs, self.t_min, self.t_max, self.hit, self.near, float(far_plane),
                        float(render_step_size), float(cone_angle), self.s_ray, self.s_cnt, self.s_ts, self.s_te,
                        aabb_p, self.s_x, self.keep_flag, int(s_cap), self.counters, self.occ_bits))
                    steps.append(PreparedCall(
                        "apnerf_field_forward_fused", self.counters[2:3], s_cap // 128, self.s_ray, self.s_cnt,
                        self.s_ts, self.s_te, self.s_x, rays_d, aabb_p, radiance_field.n_levels, meta_p, table, weights,
                        self.n_sem, state, n_rays, rays_per_call, float(alpha_thre), opc_thre, self.n_samp,
                        self.iter_samples, int(max_samples), self.keep_flag, self.total_samples,
                        1 if probabilistic else 0, ray_counts))
                    steps.append(PreparedCall(
                        "apnerf_render_compact", n_rays, rays_per_call, cur, self.keep_flag, nxt, self.n_alive_acc,
                        self.chain, self.counters))
                else:
                    steps.append(PreparedCall(
                        "apnerf_render_march", n_rays, rays_per_call, cur, self.n_samp, rays_o, rays_d, rx, ry, rz,
                        binaries, aabbs, self.t_min, self.t_max, self.hit, self.near, float(far_plane),
                        float(render_step_size), float(cone_angle), self.entry_base, self.entry_cnt, self.s_ray,
                        self.s_ts, self.s_te, aabb_p, self.s_x, self.counters, self.occ_bits))
                    steps.append(PreparedCall(
                        "apnerf_field_forward_rows", self.counters[2:3], (s_cap + 127) // 128, self.s_ray, self.s_x,
                        rays_d, aabb_p, radiance_field.n_levels, meta_p, table, weights, self.rows))
                    steps.append(PreparedCall(
                        "apnerf_render_composite", n_rays, n_rays, rays_per_call, self.n_sem, cur, self.entry_base,
                        self.entry_cnt, self.s_ts, self.s_te, self.rows, state, float(alpha_thre), opc_thre,
                        self.n_samp, self.iter_samples, int(max_samples), nxt, self.n_alive_acc, self.total_samples,
                        self.counters, 1 if probabilistic else 0, ray_counts))
                seq.append(steps)
            self._seq = seq
            for it in range(max_iters):
                steps = seq[it % 2]
                steps[0]()
                steps[1]()
                if debug_hook is not None:
                    debug_hook(it, self)
                steps[2]()
                steps[3]()
                if poll_every and (it + 1) % poll_every == 0:
                    self._pinned[it:it + 1].copy_(self.counters[1:2], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    events.append((it, ev))
                    if len(events) > ahead:  # throttle: never run more than `ahead` polls ahead of the device
                        t_wait = time.perf_counter()
                        events[-1 - ahead][1].synchronize()
                        FusedRenderer.host_blocked_s += time.perf_counter() - t_wait
                    finished = False
                    while events and events[0][1].query():
                        j, _ = events.pop(0)
                        finished = finished or int(self._pinned[j]) == 0
                    if finished:
                        break
                yield state
            seq[0][0]()  # one more schedule launch folds the last iteration's rows into counters[8]
        yield state

    def use_current_stream(self) -> None:
        """Move the REST of the render in flight (``render_iter``) to torch's current CUDA stream: the pre-marshalled
        per-iteration launches are re-targeted.  The caller orders the two streams (event record / wait)."""
        for steps in getattr(self, "_seq", ()):
            for pc in steps:
                pc.use_current_stream()

    def rows_evaluated(self) -> int:
        """Sample rows sent through the field by the last render (host sync)."""
        return int(self.counters[8].item())

    def check_overflow(self) -> None:
        """Raise if a marching iteration needed more sample rows than were allocated (host sync)."""
        if hasattr(self, "counters") and int(self.counters[5].item()) != 0:
            self.counters[5] = 0
            raise RuntimeError("apnerf fused renderer: per-iteration sample buffer overflow (results invalid)")

    @torch.no_grad()
    def finalize(self, state: Tensor, render_bkgd: Optional[Tensor] = None, want=("rgb", "rgb_var", "opacity", "depth",
                                                                                  "depth_var", "sem")):
        """utils.py:1012-1023 -> dict of [n_rays, D] tensors."""
        n_rays = state.shape[1]
        dev = state.device
        bk = [0.0, 0.0, 0.0] if render_bkgd is None else [float(v) for v in render_bkgd.reshape(-1)[:3].tolist()]
        out = {}
        shapes = dict(rgb=3, rgb_var=3, opacity=1, depth=1, depth_var=1, sem=self.n_sem)
        for k in want:
            if shapes[k] > 0:
                out[k] = torch.empty((n_rays, shapes[k]), device=dev)
        with torch.cuda.device(dev):
            call("apnerf_render_finalize", n_rays, self.n_sem, state, bk[0], bk[1], bk[2], out.get("rgb"),
                 out.get("rgb_var"), out.get("opacity"), out.get("depth"), out.get("depth_var"), out.get("sem"))
        return out


_RENDERERS: Dict[Tuple[str, int], FusedRenderer] = {}


def _renderer_for(device, n_sem) -> FusedRenderer:
    key = (str(device), int(n_sem))
    if key not in _RENDERERS:
        _RENDERERS[key] = FusedRenderer(device, n_sem)
    return _RENDERERS[key]


def _flatten_rays(rays: Rays):
    rays_shape = rays.origins.shape
    if len(rays_shape) == 3:
        height, width, _ = rays_shape
        num_rays = height * width
        rays = namedtuple_map(lambda r: r.reshape([num_rays] + list(r.shape[2:])), rays)
    else:
        num_rays, _ = rays_shape
    return rays, rays_shape, num_rays


@torch.no_grad()
def render_probablistic_image_with_occgrid_test(
    max_samples: int, radiance_field: torch.nn.Module, estimator: OccGridEstimator, rays: Rays,
    near_plane: float = 0.0, far_plane: float = 1e10, render_step_size: float = 1e-3,
    render_bkgd: Optional[torch.Tensor] = None, cone_angle: float = 0.0, alpha_thre: float = 0.0,
    early_stop_eps: float = 1e-4, timestamps: Optional[torch.Tensor] = None,
):
    """Drop-in for perception/models/utils.py:782-1032.  Returns
    (rgb, rgb_var, opacity, depth, depth_var[, sem], total_samples)."""
    assert timestamps is None, "dnerf timestamps are not part of the pipeline's path"
    if estimator.binaries.shape[0] != 1:
        return render_probablistic_image_with_occgrid_test_unfused(
            max_samples, radiance_field, estimator, rays, near_plane, far_plane, render_step_size, render_bkgd,
            cone_angle, alpha_thre, early_stop_eps)
    rays, rays_shape, num_rays = _flatten_rays(rays)
    C = radiance_field.num_semantic_classes
    r = _renderer_for(rays.origins.device, C)
    state = r.render(radiance_field, estimator, rays.origins, rays.viewdirs, num_rays, max_samples=max_samples,
                     near_plane=near_plane, far_plane=far_plane, render_step_size=render_step_size,
                     cone_angle=cone_angle, alpha_thre=alpha_thre, early_stop_eps=early_stop_eps, probabilistic=True)
    o = r.finalize(state, render_bkgd)
    total = int(r.total_samples[0].item())
    r.check_overflow()
    view = lambda t: t.view((*rays_shape[:-1], -1))
    if C > 0:
        return (view(o["rgb"]), view(o["rgb_var"]), view(o["opacity"]), view(o["depth"]), view(o["depth_var"]),
                view(o["sem"]), total)
    return view(o["rgb"]), view(o["rgb_var"]), view(o["opacity"]), view(o["depth"]), view(o["depth_var"]), total


@torch.no_grad()
def render_image_with_occgrid_test(
    max_samples: int, radiance_field: torch.nn.Module, estimator: OccGridEstimator, rays: Rays,
    near_plane: float = 0.0, far_plane: float = 1e10, render_step_size: float = 1e-3,
    render_bkgd: Optional[torch.Tensor] = None, cone_angle: float = 0.0, alpha_thre: float = 0.0,
    early_stop_eps: float = 1e-4, timestamps: Optional[torch.Tensor] = None,
):
    """Drop-in for perception/models/utils.py:555-779: (rgb, opacity, depth[, sem], total_samples)."""
    assert timestamps is None, "dnerf timestamps are not part of the pipeline's path"
    if estimator.binaries.shape[0] != 1:
        out = render_probablistic_image_with_occgrid_test_unfused(
            max_samples, radiance_field, estimator, rays, near_plane, far_plane, render_step_size, render_bkgd,
            cone_angle, alpha_thre, early_stop_eps, probabilistic=False)
        return out
    rays, rays_shape, num_rays = _flatten_rays(rays)
    C = radiance_field.num_semantic_classes
    r = _renderer_for(rays.origins.device, C)
    state = r.render(radiance_field, estimator, rays.origins, rays.viewdirs, num_rays, max_samples=max_samples,
                     near_plane=near_plane, far_plane=far_plane, render_step_size=render_step_size,
                     cone_angle=cone_angle, alpha_thre=alpha_thre, early_stop_eps=early_stop_eps, probabilistic=False)
    o = r.finalize(state, render_bkgd, want=("rgb", "opacity", "depth", "sem"))
    total = int(r.total_samples[0].item())
    r.check_overflow()
    view = lambda t: t.view((*rays_shape[:-1], -1))
    if C > 0:
        return view(o["rgb"]), view(o["opacity"]), view(o["depth"]), view(o["sem"]), total
    return view(o["rgb"]), view(o["opacity"]), view(o["depth"]), total


# ------------------------------------------------------------------------------------------
# Op-by-op version (reference control flow on the drop-in ops)
# ------------------------------------------------------------------------------------------
@torch.no_grad()
def render_probablistic_image_with_occgrid_test_unfused(
    max_samples: int, radiance_field: torch.nn.Module, estimator: OccGridEstimator, rays: Rays,
    near_plane: float = 0.0, far_plane: float = 1e10, render_step_size: float = 1e-3,
    render_bkgd: Optional[torch.Tensor] = None, cone_angle: float = 0.0, alpha_thre: float = 0.0,
    early_stop_eps: float = 1e-4, probabilistic: bool = True, trace: Optional[list] = None,
):
    rays, rays_shape, num_rays = _flatten_rays(rays)
    rays_o, rays_d = rays.origins, rays.viewdirs
    device = rays_o.device
    C = radiance_field.num_semantic_classes
    opacity = torch.zeros(num_rays, 1, device=device)
    depth = torch.zeros(num_rays, 1, device=device)
    rgb = torch.zeros(num_rays, 3, device=device)
    sem = torch.zeros(num_rays, C, device=device)
    depth_var = torch.zeros(num_rays, 1, device=device)
    rgb_var = torch.zeros(num_rays, 3, device=device)
    ray_mask = torch.ones(num_rays, device=device).bool()
    min_samples = 1 if cone_angle == 0 else 4
    iter_samples = total_samples = 0
    near_planes = torch.full_like(rays_o[..., 0], fill_value=near_plane)
    far_planes = torch.full_like(rays_o[..., 0], fill_value=far_plane)
    t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, estimator.aabbs)
    n_grids = estimator.binaries.size(0)
    if n_grids > 1:
        t_sorted, t_indices = torch.sort(torch.cat([t_mins, t_maxs], -1), -1)
    else:
        t_sorted = torch.cat([t_mins, t_maxs], -1)
        t_indices = torch.arange(0, n_grids * 2, device=device, dtype=torch.int64).expand(num_rays, n_grids * 2)
    opc_thre = 1 - early_stop_eps
    if render_bkgd is None:
        render_bkgd = torch.zeros(3, device=device)
    while iter_samples < max_samples:
        n_alive = ray_mask.sum().item()
        if n_alive == 0:
            break
        n_samples = max(min(num_rays // n_alive, 64), min_samples)
        iter_samples += n_samples
        intervals, samples, termination_planes = traverse_grids(
            rays_o, rays_d, estimator.binaries, estimator.aabbs, near_planes, far_planes, render_step_size,
            cone_angle, n_samples, True, ray_mask, t_sorted, t_indices, hits)
        t_starts = intervals.vals[intervals.is_left]
        t_ends = intervals.vals[intervals.is_right]
        ray_indices = samples.ray_indices[samples.is_valid]
        packed_info = samples.packed_info
        if trace is not None:
            trace.append(dict(n_alive=n_alive, n_samples=n_samples, ray_indices=ray_indices.clone(),
                              t_starts=t_starts.clone(), t_ends=t_ends.clone()))
        t_dirs = rays_d[ray_indices]
        positions = rays_o[ray_indices] + t_dirs * (t_starts[:, None] + t_ends[:, None]) / 2.0
        if positions.shape[0] == 0:
            rgbs = torch.zeros(0, 3, device=device)
            sigmas = torch.zeros(0, device=device)
            sems = torch.zeros(0, C, device=device)
        elif C > 0:
            rgbs, sigmas, sems = radiance_field(positions, t_dirs)
            sigmas = sigmas.squeeze(-1)
        else:
            rgbs, sigmas = radiance_field(positions, t_dirs)
            sigmas, sems = sigmas.squeeze(-1), torch.zeros(positions.shape[0], 0, device=device)
        weights, _, alphas = render_weight_from_density(
            t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=num_rays,
            prefix_trans=1 - opacity[ray_indices].squeeze(-1))
        if alpha_thre > 0:
            vis = alphas >= alpha_thre
            ray_indices, rgbs, weights, t_starts, t_ends, sems = (
                ray_indices[vis], rgbs[vis], weights[vis], t_starts[vis], t_ends[vis], sems[vis])
        t_mid = (t_starts + t_ends)[..., None] / 2.0
        accumulate_along_rays_(weights, values=rgbs, ray_indices=ray_indices, outputs=rgb)
        accumulate_along_rays_(weights, values=None, ray_indices=ray_indices, outputs=opacity)
        accumulate_along_rays_(weights, values=t_mid, ray_indices=ray_indices, outputs=depth)
        if C > 0:
            accumulate_along_rays_(weights, values=sems, ray_indices=ray_indices, outputs=sem)
        if probabilistic:
            accumulate_along_rays_(weights, values=torch.pow(rgbs - rgb[ray_indices], 2), ray_indices=ray_indices,
                                   outputs=rgb_var)
            accumulate_along_rays_(weights, values=torch.pow(t_mid - depth[ray_indices], 2), ray_indices=ray_indices,
                                   outputs=depth_var)
        # masked rays keep their previous terminate plane (the reference leaves them uninitialised
        # but never marches them again)
        near_planes = torch.where(ray_mask, termination_planes, near_planes)
        ray_mask = torch.logical_and(opacity.view(-1) <= opc_thre, packed_info[:, 1] == n_samples)
        total_samples += ray_indices.shape[0]
    rgb = rgb + render_bkgd * (1.0 - opacity)
    depth = depth / opacity.clamp_min(torch.finfo(rgb.dtype).eps)
    view = lambda t: t.view((*rays_shape[:-1], -1))
    if probabilistic:
        if C > 0:
            return view(rgb), view(rgb_var), view(opacity), view(depth), view(depth_var), view(sem), total_samples
        return view(rgb), view(rgb_var), view(opacity), view(depth), view(depth_var), total_samples
    if C > 0:
        return view(rgb), view(opacity), view(depth), view(sem), total_samples
    return view(rgb), view(opacity), view(depth), total_samples


# ------------------------------------------------------------------------------------------
# Train-mode glue
# ------------------------------------------------------------------------------------------
def sem_rendering(
    t_starts: Tensor, t_ends: Tensor, ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
    rgb_sigma_sem_fn: Optional[Callable] = None, render_bkgd: Optional[Tensor] = None,
    num_sumantic_classes: int = 0,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Dict]:
    """utils.py:362-461: colours, opacities, depths, semantics (+ extras), differentiable w.r.t.
    the field outputs."""
    if ray_indices is not None:
        assert t_starts.shape == t_ends.shape == ray_indices.shape
    if rgb_sigma_sem_fn is None:
        raise ValueError("At least one of `rgb_sigma_fn` and `rgb_alpha_fn` should be specified.")
    if t_starts.shape[0] != 0:
        rgbs, sigmas, sems = rgb_sigma_sem_fn(t_starts, t_ends, ray_indices)
    else:
        rgbs = torch.empty((0, 3), device=t_starts.device)
        sigmas = torch.empty((0,), device=t_starts.device)
        sems = torch.empty((0, num_sumantic_classes), device=t_starts.device)
    assert rgbs.shape[-1] == 3, "rgbs must have 3 channels, got {}".format(rgbs.shape)
    assert sigmas.shape == t_starts.shape, "sigmas must have shape of (N,)! Got {}".format(sigmas.shape)
    assert sems.shape[-1] == num_sumantic_classes
    weights, trans, alphas = render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices,
                                                        n_rays=n_rays)
    extras = {"weights": weights, "alphas": alphas, "trans": trans, "sigmas": sigmas, "rgbs": rgbs}
    colors = accumulate_along_rays(weights, values=rgbs, ray_indices=ray_indices, n_rays=n_rays)
    opacities = accumulate_along_rays(weights, values=None, ray_indices=ray_indices, n_rays=n_rays)
    depths = accumulate_along_rays(weights, values=(t_starts + t_ends)[..., None] / 2.0, ray_indices=ray_indices,
                                   n_rays=n_rays)
    depths = depths / opacities.clamp_min(torch.finfo(rgbs.dtype).eps)
    semantics = accumulate_along_rays(weights, values=sems, ray_indices=ray_indices, n_rays=n_rays)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opacities)
    return colors, opacities, depths, semantics, extras


def render_image_with_occgrid(
    radiance_field: torch.nn.Module, estimator: OccGridEstimator, rays: Rays, near_plane: float = 0.0,
    far_plane: float = 1e10, render_step_size: float = 1e-3, render_bkgd: Optional[torch.Tensor] = None,
    cone_angle: float = 0.0, alpha_thre: float = 0.0, test_chunk_size: int = 8192,
    timestamps: Optional[torch.Tensor] = None,
):
    """Drop-in for perception/models/utils.py:222-359 (the depth-guided variant below without the guide)."""
    return render_image_with_occgrid_with_depth_guide(
        radiance_field, estimator, rays, near_plane=near_plane, far_plane=far_plane, render_step_size=render_step_size,
        render_bkgd=render_bkgd, cone_angle=cone_angle, alpha_thre=alpha_thre, test_chunk_size=test_chunk_size,
        timestamps=timestamps, depth=None)


def render_image_with_occgrid_with_depth_guide(
    radiance_field: torch.nn.Module, estimator: OccGridEstimator, rays: Rays, near_plane: float = 0.0,
    far_plane: float = 1e10, render_step_size: float = 1e-3, render_bkgd: Optional[torch.Tensor] = None,
    cone_angle: float = 0.0, alpha_thre: float = 0.0, test_chunk_size: int = 8192,
    timestamps: Optional[torch.Tensor] = None, depth: Optional[torch.Tensor] = None,
):
    """Train-mode render, drop-in for perception/models/utils.py:63-219: occupancy-grid sampling with a
    no-grad density pre-filter (stratified when training), then a differentiable field query on the
    surviving samples and packed compositing.  Returns (rgb, opacity, depth[, semantics], n_samples)."""
    assert timestamps is None, "dnerf timestamps are not part of the pipeline's path"
    rays, rays_shape, num_rays = _flatten_rays(rays)
    C = radiance_field.num_semantic_classes
    results = []
    chunk = torch.iinfo(torch.int32).max if radiance_field.training else test_chunk_size
    for i in range(0, num_rays, chunk):
        chunk_rays = namedtuple_map(lambda r: r[i:i + chunk], rays)

        def positions_of(t_starts, t_ends, ray_indices):
            t_dirs = chunk_rays.viewdirs[ray_indices]
            return chunk_rays.origins[ray_indices] + t_dirs * (t_starts + t_ends)[:, None] / 2.0, t_dirs

        def sigma_fn(t_starts, t_ends, ray_indices):
            return radiance_field.query_density(positions_of(t_starts, t_ends, ray_indices)[0]).squeeze(-1)

        def rgb_sigma_sem_fn(t_starts, t_ends, ray_indices):
            positions, t_dirs = positions_of(t_starts, t_ends, ray_indices)
            out = radiance_field(positions, t_dirs)
            return (out[0], out[1].squeeze(-1)) + tuple(out[2:])

        ray_indices, t_starts, t_ends = estimator.sampling(
            chunk_rays.origins, chunk_rays.viewdirs, sigma_fn=sigma_fn, near_plane=near_plane, far_plane=far_plane,
            render_step_size=render_step_size, stratified=radiance_field.training, cone_angle=cone_angle,
            alpha_thre=alpha_thre, depth=depth)
        n_chunk = chunk_rays.origins.shape[0]
        if C > 0:
            rgb, opacity, dep, semantics, _ = sem_rendering(
                t_starts, t_ends, ray_indices, n_rays=n_chunk, rgb_sigma_sem_fn=rgb_sigma_sem_fn,
                render_bkgd=render_bkgd, num_sumantic_classes=C)
            results.append([rgb, opacity, dep, semantics, len(t_starts)])
        else:
            from .nerfacc import rendering

            rgb, opacity, dep, _ = rendering(t_starts, t_ends, ray_indices, n_rays=n_chunk,
                                             rgb_sigma_fn=rgb_sigma_sem_fn, render_bkgd=render_bkgd)
            results.append([rgb, opacity, dep, len(t_starts)])
    cols = [torch.cat(r, dim=0) if isinstance(r[0], torch.Tensor) else sum(r) for r in zip(*results)]
    return tuple(c.view((*rays_shape[:-1], -1)) if isinstance(c, torch.Tensor) else c for c in cols)
