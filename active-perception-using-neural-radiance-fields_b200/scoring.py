"""Candidate-trajectory scoring: render the planner's candidate views through every ensemble
member and reduce them to predictive information -- the caller side of the hot path.

Mirrors, on the reference side,
  ActiveNeRFMapper.probablistic_uncertainty          scripts/pipeline.py:666-798
  Dataset.render_probablistic_image_from_pose        perception/data_proc/habitat_to_data.py:413-549
with the renders kept on the device (the reference copies six arrays per view to host numpy and
reduces them in float64 numpy) and the planner's loop over trajectories (pipeline.py:1079-1085)
folded into one batch.  Multi-GPU: views are sharded over ranks, each rank reduces its views to
per-trajectory partial sums, and ONE all-reduce of [n_traj, 4] float64 finishes the job.
"""
import collections
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import call
from .render import FusedRenderer


def quat_xyzw_to_matrix(q) -> np.ndarray:
    """Rotation matrix of a (x, y, z, w) quaternion, as scipy's Rotation.from_quat(q).as_matrix()
    (habitat_to_data.py:445-449)."""
    x, y, z, w = (float(v) for v in q)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def poses_to_c2w(poses: np.ndarray) -> np.ndarray:
    """[n, 7] (x, y, z, qx, qy, qz, qw) float64 -> [n, 3, 4] float32 camera-to-world."""
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 7)
    out = np.empty((poses.shape[0], 3, 4), dtype=np.float32)
    for i, p in enumerate(poses):
        out[i, :, :3] = quat_xyzw_to_matrix(p[3:])
        out[i, :, 3] = p[:3]
    return out


def uncertainty_view_indices(traj_len: int) -> np.ndarray:
    """The 40 views the reference scores per trajectory (pipeline.py:687-689)."""
    a = np.linspace(0, traj_len - 20, 20)
    b = np.linspace(traj_len - 20, traj_len - 1, 20)
    return np.hstack((a, b)).astype(int)


class PredictiveInformationScorer:
    """Renders views x ensemble members with the fused renderer and scores them on the device."""

    def __init__(self, radiance_fields: Sequence[torch.nn.Module], estimators: Sequence[torch.nn.Module], width: int,
                 height: int, focal: float, *, near_plane: float = 0.1, render_step_size: float = 1e-3,
                 cone_angle: float = 0.004, alpha_thre: float = 0.01, scale: float = 1.0, max_samples: int = 1024,
                 device="cuda:0", views_per_batch: int = 32, concurrent_batches: int = 1, balance: str = "lpt"):
        assert 1 <= len(radiance_fields) <= 4 and len(radiance_fields) == len(estimators)
        self.fields, self.estimators = list(radiance_fields), list(estimators)
        self.width, self.height, self.focal = int(width), int(height), float(focal)
        self.opts = dict(near_plane=near_plane, render_step_size=render_step_size, cone_angle=cone_angle,
                         alpha_thre=alpha_thre, max_samples=max_samples)
        self.device = torch.device(device)
        self.n_sem = self.fields[0].num_semantic_classes
        self.views_per_batch = int(views_per_batch)
        assert balance in ("lpt", "contiguous")
        self.balance = balance
        self.after_render = None  # optional callable(renderer), invoked once per finished render (measurement)
        self._probe = None
        # rounded-linspace subsample of the full image (habitat_to_data.py:462-467)
        h, w = int(height * scale), int(width * scale)
        self.rays_per_view = h * w
        if self.rays_per_view == width * height:
            self.keep_idx = None
        else:
            idx = np.round(np.linspace(0, width * height - 1, self.rays_per_view)).astype(np.int32)
            self.keep_idx = torch.from_numpy(idx).to(self.device)
        # Renders in flight: `concurrent_batches` view batches x the ensemble members, each with its own
        # working set and stream.  The marching loops are independent, so the narrow, launch-bound tail
        # iterations of one render overlap the wide early iterations of the others.
        self.concurrent_batches = max(1, int(concurrent_batches))
        self.renderers = [[FusedRenderer(self.device, self.n_sem) for _ in self.fields]
                          for _ in range(self.concurrent_batches)]
        self.renderer = self.renderers[0][0]
        self._streams = None
        self.interleave = True  # False: one render after the other on the current stream (measurement)
        self.stagger_iters = int(os.environ.get("APNERF_STAGGER", "12"))  # a batch leaves its head phase after this many marching iterations (see partial_sums)
        self._states = None
        self._rays = None

    def all_renderers(self):
        return [r for slot in self.renderers for r in slot]

    def _buffers(self, n_views):
        n_rays = n_views * self.rays_per_view
        if self._rays is None or self._rays[0][0].shape[0] < n_rays:
            self._rays = [(torch.empty((n_rays, 3), device=self.device), torch.empty((n_rays, 3), device=self.device))
                          for _ in range(self.concurrent_batches)]
            self._states = [[torch.empty((9 + self.n_sem, n_rays), device=self.device) for _ in self.fields]
                            for _ in range(self.concurrent_batches)]
        return n_rays

    @torch.no_grad()
    def partial_sums(self, c2w: torch.Tensor, view_traj: torch.Tensor, n_traj: int,
                     sums: Optional[torch.Tensor] = None) -> torch.Tensor:
        """c2w [n_views, 3, 4] f32 and view_traj [n_views] i32 ON THE DEVICE -> float64 [n_traj, 4]
        sums of the per-pixel (rgb, depth, sem, occ) predictive-information terms.  Everything is
        enqueued on CUDA streams; nothing is read back except the renderers' non-blocking look at
        their live-ray counters."""
        n_views = c2w.shape[0]
        if sums is None:
            sums = torch.zeros((n_traj, 4), device=self.device, dtype=torch.float64)
        vb = self.views_per_batch
        self._buffers(min(vb, n_views))
        E, K = len(self.fields), self.concurrent_batches
        with torch.cuda.device(self.device):
            if self._streams is None:
                self._streams = [[torch.cuda.Stream(device=self.device) for _ in range(E)] for _ in range(K)]
            main = torch.cuda.current_stream()
            batches = collections.deque((v0, min(n_views, v0 + vb)) for v0 in range(0, n_views, vb))
            free_slots = list(range(K))
            active = []  # batches in flight: dict(slot, v0, v1, nr, states, gens = [[stream, generator, iterations]])

            def start(v0, v1, slot):
                nr = (v1 - v0) * self.rays_per_view
                rays_o, rays_d = self._rays[slot][0][:nr], self._rays[slot][1][:nr]
                call("apnerf_generate_rays", v1 - v0, c2w[v0:v1].contiguous(), self.width, self.height, self.focal,
                     self.rays_per_view, self.keep_idx, rays_o, rays_d)
                ready = torch.cuda.Event()
                ready.record(main)
                states, gens = [], []
                for m, (f, e) in enumerate(zip(self.fields, self.estimators)):
                    st = self._states[slot][m].view(-1)[: (9 + self.n_sem) * nr].view(9 + self.n_sem, nr)
                    states.append(st)
                    r = self.renderers[slot][m]
                    if e.binaries.shape[0] != 1:  # multi-level grids: the op-by-op renderer, view by view
                        self._render_unfused(f, e, rays_o, rays_d, st)
                    elif not self.interleave:
                        r.render(f, e, rays_o, rays_d, self.rays_per_view, probabilistic=True, state=st, **self.opts)
                        if self.after_render is not None:
                            self.after_render(r)
                    else:
                        self._streams[slot][m].wait_event(ready)
                        gens.append([self._streams[slot][m],
                                     r.render_iter(f, e, rays_o, rays_d, self.rays_per_view, probabilistic=True, state=st,
                                                   **self.opts), 0])
                return dict(slot=slot, v0=v0, v1=v1, nr=nr, states=states, gens=gens)

            def finish(job):
                for stream, _, _ in job["gens"]:
                    done = torch.cuda.Event()
                    done.record(stream)
                    main.wait_event(done)
                if self.after_render is not None and self.interleave:
                    for r in self.renderers[job["slot"]]:
                        self.after_render(r)
                states = job["states"] + [None] * (4 - len(job["states"]))
                call("apnerf_score_views", E, states[0], states[1], states[2], states[3], job["nr"], self.rays_per_view,
                     self.n_sem, view_traj[job["v0"]:job["v1"]].contiguous(), n_traj, sums)
                free_slots.append(job["slot"])

            # Rolling pipeline.  The first dozen marching iterations of a batch are wide, throughput-bound launches; the
            # long tail (a few views whose rays cross much transparent occupied space keep marching 4 samples at a time,
            # up to 256 iterations) is a train of small latency-bound launches that leave most of the GPU idle.  A new
            # batch is therefore started, on its own streams, as soon as every batch in flight has left its head phase:
            # the next head's big kernels fill the SMs the tails do not use.
            while batches or active:
                head_done = all(g[2] >= self.stagger_iters or g[1] is None for job in active for g in job["gens"])
                if batches and free_slots and (not active or head_done):
                    v0, v1 = batches.popleft()
                    active.append(start(v0, v1, free_slots.pop(0)))
                for job in list(active):
                    running = False
                    for g in job["gens"]:  # one marching iteration per render, each on its own stream
                        if g[1] is None:
                            continue
                        with torch.cuda.stream(g[0]):
                            if next(g[1], None) is None:
                                g[1] = None
                            else:
                                g[2] += 1
                                running = True
                    if not running:
                        active.remove(job)
                        finish(job)
        return sums

    def _render_unfused(self, field, estimator, rays_o, rays_d, st):
        """Fill the state planes the scorer reads (opacity, variances, semantic logits) from the op-by-op renderer,
        one call per view as the reference does: the path for occupancy grids with more than one level."""
        from .render import Rays, render_probablistic_image_with_occgrid_test_unfused

        R = self.rays_per_view
        for v0 in range(0, rays_o.shape[0], R):
            rays = Rays(origins=rays_o[v0:v0 + R], viewdirs=rays_d[v0:v0 + R])
            out = render_probablistic_image_with_occgrid_test_unfused(
                self.opts["max_samples"], field, estimator, rays, near_plane=self.opts["near_plane"],
                render_step_size=self.opts["render_step_size"], cone_angle=self.opts["cone_angle"],
                alpha_thre=self.opts["alpha_thre"])
            rgb_var, opacity, depth_var = out[1], out[2], out[4]
            st[5:8, v0:v0 + R] = rgb_var.t()
            st[8, v0:v0 + R] = depth_var[:, 0]
            st[3, v0:v0 + R] = opacity[:, 0]
            if self.n_sem > 0:
                st[9:, v0:v0 + R] = out[5].t()

    # ---- multi-GPU view assignment ---------------------------------------------------------------------------
    @torch.no_grad()
    def estimate_rows(self, c2w: torch.Tensor) -> torch.Tensor:
        """Estimated field evaluations per view ([n_views] float32, device): the views are rendered through member 0 at
        1/64 of their rays (rounded-linspace subsample, the reference's own subsampling rule) and the sample rows
        the marcher emitted are counted per view.  On the synthetic scene the estimate tracks the full-resolution
        count to a few per cent while a view's cost varies 3x between poses."""
        n_views = c2w.shape[0]
        if self._probe is None:
            k = int(min(self.rays_per_view, max(256, self.rays_per_view // 64)))
            idx = np.round(np.linspace(0, self.rays_per_view - 1, k)).astype(np.int64)
            if self.keep_idx is not None:
                idx = self.keep_idx.cpu().numpy().astype(np.int64)[idx]
            self._probe = dict(k=k, keep=torch.from_numpy(idx.astype(np.int32)).to(self.device),
                               renderer=FusedRenderer(self.device, self.n_sem))
        pr = self._probe
        k = pr["k"]
        n_rays = n_views * k
        if pr.get("cap", 0) < n_rays:
            pr["rays_o"] = torch.empty((n_rays, 3), device=self.device)
            pr["rays_d"] = torch.empty((n_rays, 3), device=self.device)
            pr["counts"] = torch.empty((2, n_rays), device=self.device, dtype=torch.int32)
            pr["cap"] = n_rays
        rays_o, rays_d = pr["rays_o"][:n_rays], pr["rays_d"][:n_rays]
        counts = pr["counts"].view(-1)[: 2 * n_rays].view(2, n_rays)
        counts.zero_()
        with torch.cuda.device(self.device):
            call("apnerf_generate_rays", n_views, c2w.contiguous(), self.width, self.height, self.focal, k, pr["keep"],
                 rays_o, rays_d)
            pr["renderer"].render(self.fields[0], self.estimators[0], rays_o, rays_d, k, probabilistic=False,
                                  ray_counts=counts, **self.opts)
        return counts[0].view(n_views, k).sum(1).float() * (self.rays_per_view / k)

    def assign_views(self, poses: np.ndarray, rank: int, world: int, process_group=None) -> np.ndarray:
        """Indices (ascending) of the views this rank renders.  world == 1 or balance == "contiguous": the contiguous
        balanced slice.  balance == "lpt": every rank estimates the cost of its contiguous slice (``estimate_rows``),
        the estimates are all-gathered (one small collective) and the views are dealt out longest-processing-time
        first to the least loaded rank -- the same deterministic assignment on every rank."""
        import torch.distributed as dist

        n = len(poses)
        lo, hi = shard_range(n, rank, world)
        if (world == 1 or self.balance == "contiguous" or n < 2 * world
                or any(e.binaries.shape[0] != 1 for e in self.estimators)):
            return np.arange(lo, hi)
        cap = (n + world - 1) // world
        est = torch.zeros(cap, device=self.device)
        if hi > lo:
            c2w = torch.from_numpy(poses_to_c2w(poses[lo:hi])).to(self.device)
            est[: hi - lo] = self.estimate_rows(c2w)
        gathered = [torch.empty_like(est) for _ in range(world)]
        dist.all_gather(gathered, est, group=process_group)
        g = torch.stack(gathered).cpu().numpy()
        cost = np.concatenate([g[r, : shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r in range(world)])
        return lpt_assign(cost, world)[rank]

    @staticmethod
    def finish(sums: np.ndarray, pixels_per_traj: np.ndarray) -> np.ndarray:
        """[n_traj, 4] sums + pixel counts -> the four entries the reference logs per trajectory
        (rgb, depth, 3 * sem, 2 * occ; pipeline.py:772-790).  Their row sum is the score."""
        sums = np.asarray(sums, dtype=np.float64)
        n = np.maximum(np.asarray(pixels_per_traj, dtype=np.float64), 1.0)[:, None]
        terms = sums / n
        terms[:, 0] /= 3.0  # the rgb mean runs over pixels x 3 channels
        terms[:, 2] *= 3.0
        terms[:, 3] *= 2.0
        return terms

    @torch.no_grad()
    def score_trajectories(self, trajectories: List[np.ndarray], process_group=None) -> np.ndarray:
        """Host-facing call: list of planner trajectories ([len, 7] pose arrays) -> [n_traj, 4]
        predictive-information terms.  With torch.distributed initialised (or a process group
        given) the views are sharded contiguously over the ranks."""
        import torch.distributed as dist

        poses, owner = [], []
        for t, traj in enumerate(trajectories):
            traj = np.asarray(traj)
            # always the reference's 40 indices: a short trajectory repeats views (and, below 20 poses, numpy's
            # negative indices wrap) exactly as trajectory[unc_idx] does at pipeline.py:687-697
            idx = uncertainty_view_indices(len(traj))
            poses.append(traj[idx])
            owner += [t] * len(idx)
        poses = np.concatenate(poses, 0)
        owner = np.asarray(owner, dtype=np.int32)
        return self.score_views(poses, owner, len(trajectories), process_group=process_group)

    @torch.no_grad()
    def score_views(self, poses: np.ndarray, view_traj: np.ndarray, n_traj: int, process_group=None) -> np.ndarray:
        import torch.distributed as dist

        use_dist = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank(process_group) if use_dist else 0
        world = dist.get_world_size(process_group) if use_dist else 1
        n_views = poses.shape[0]
        mine = self.assign_views(poses, rank, world, process_group)
        c2w_host = torch.from_numpy(poses_to_c2w(poses[mine])).pin_memory()
        vt_host = torch.from_numpy(np.ascontiguousarray(np.asarray(view_traj)[mine], dtype=np.int32)).pin_memory()
        c2w = c2w_host.to(self.device, non_blocking=True)
        vt = vt_host.to(self.device, non_blocking=True)
        sums = torch.zeros((n_traj, 4), device=self.device, dtype=torch.float64)
        if len(mine):
            self.partial_sums(c2w, vt, n_traj, sums)
        sums = all_reduce_partial_sums(sums, process_group)
        counts = np.bincount(view_traj, minlength=n_traj)[:n_traj] * self.rays_per_view
        host_sums = sums.cpu().numpy()
        for r in self.all_renderers():
            r.check_overflow()
        return self.finish(host_sums, counts)


def all_reduce_partial_sums(sums: torch.Tensor, process_group=None) -> torch.Tensor:
    """The ONE collective of the render + score path: sum the per-rank [n_traj, 4] float64 partial
    sums (NCCL over NVLink on GPUs; any torch.distributed backend works)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=process_group)
    return sums


def lpt_assign(cost: np.ndarray, world: int):
    """Longest-processing-time-first assignment of units with the given costs to `world` bins; returns one ascending
    index array per bin.  Deterministic (stable sort, ties to the lower bin), so every rank computes the same split."""
    order = np.argsort(-np.asarray(cost, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    bins = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += max(float(cost[i]), 0.0) + 1e-9
    return [np.asarray(sorted(b), dtype=np.int64) for b in bins]


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of n units for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def probablistic_uncertainty(radiance_fields, estimators, trajectory, *, img_w, img_h, focal, near_plane,
                             render_step_size, cone_angle, alpha_thre, scale=0.1, device="cuda:0", log=None) -> float:
    """Drop-in for the body of ActiveNeRFMapper.probablistic_uncertainty (pipeline.py:666-798):
    returns the trajectory's predictive information; appends the four logged terms to `log`."""
    key = (tuple(id(f) for f in radiance_fields), tuple(id(e) for e in estimators), img_w, img_h, float(focal),
           float(scale), str(device), float(near_plane), float(render_step_size), float(cone_angle), float(alpha_thre))
    scorer = _SCORERS.get(key)
    if scorer is None:
        scorer = PredictiveInformationScorer(radiance_fields, estimators, img_w, img_h, focal, near_plane=near_plane,
                                             render_step_size=render_step_size, cone_angle=cone_angle,
                                             alpha_thre=alpha_thre, scale=scale, device=device, views_per_batch=40)
        _SCORERS.clear()
        _SCORERS[key] = scorer
    terms = scorer.score_trajectories([np.asarray(trajectory)])[0]
    if log is not None:
        log.append(terms.tolist())
    return float(terms.sum())


_SCORERS = {}


def trajector_uncertainty(radiance_fields, estimators, trajectory, step, *, img_w, img_h, focal, near_plane,
                          render_step_size, cone_angle, alpha_thre, scale=0.1, device="cuda:0", log=None):
    """Drop-in for the body of the older scorer ActiveNeRFMapper.trajector_uncertainty (pipeline.py:800-916),
    used by the "random" policy: ensemble variance of rgb / depth, inverse opacity and (first member only)
    semantic entropy per view, clipped and summed.  Returns ``(uncertainty, max_idx)``; appends the logged
    per-view terms to ``log``.  The renders come from the device-driven renderer in one batch; the reduction is
    the reference's float64 numpy, line for line (it is tiny: 40 views of 1 % resolution).

    Reference quirk kept out: its tuple unpacking (4 names for member 0 at :823, 3 names for the others at
    :842) only works for ONE member with semantic classes and raises otherwise; here every ensemble shape
    is accepted and only the first member's semantics are used, as the reference intends."""
    from .data_proc import Dataset

    num_sample = 40
    trajectory = np.asarray(trajectory)
    unc_idx = uncertainty_view_indices(len(trajectory))
    rendered, depths, accs, sems = [], [], [], []
    for m, (f, e) in enumerate(zip(radiance_fields, estimators)):
        out = Dataset.render_image_from_pose(f, e, trajectory[unc_idx], img_w, img_h, focal, near_plane,
                                             render_step_size, scale, cone_angle, alpha_thre, 4, device)
        rendered.append(out[0][-num_sample:])
        depths.append(out[1][-num_sample:])
        accs.append(out[2][-num_sample:])
        if m == 0 and len(out) > 3:
            sems.append(out[3][-num_sample:])
    rendered, depths = np.array(rendered), np.array(depths)
    acc0 = np.array(accs[0]) + 1e-4
    intensity_var = np.mean(np.var(rendered, axis=0), axis=-1)
    depth_var = np.var(depths, axis=0)
    intensity_var_mean = np.clip(np.mean(intensity_var, axis=(1, 2)) * 4000, 0, 100)
    depth_var_mean = np.clip(np.mean(depth_var, axis=(1, 2)) * 50, 0, 100)
    acc_inv_mean = np.mean(np.clip(1 / acc0 - 1, 0, 10000), axis=(1, 2))
    terms = [intensity_var_mean[-num_sample:], depth_var_mean[-num_sample:], acc_inv_mean[-num_sample:]]
    uncertainty = intensity_var_mean + depth_var_mean + acc_inv_mean
    if radiance_fields[0].num_semantic_classes > 0:
        sem_p = torch.softmax(torch.from_numpy(np.array(sems)), dim=-1).numpy()
        sem_entropy = -np.sum(sem_p * np.log(sem_p + 1e-10), axis=-1)
        sem_entropy_mean = np.clip(np.mean(sem_entropy, axis=(0, 2, 3)) * 50, 0, 100)
        uncertainty = uncertainty + sem_entropy_mean
        terms.append(sem_entropy_mean[-num_sample:])
    max_idx = np.sort(np.argsort(uncertainty))
    uncertainty = np.mean(uncertainty[-11:]) if step == -1 else np.mean(uncertainty[max_idx])
    if log is not None:
        log.append(terms)
    return uncertainty, max_idx
