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
import time
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
    _instances = 0

    def __init__(self, radiance_fields: Sequence[torch.nn.Module], estimators: Sequence[torch.nn.Module], width: int,
                 height: int, focal: float, *, near_plane: float = 0.1, render_step_size: float = 1e-3,
                 cone_angle: float = 0.004, alpha_thre: float = 0.01, scale: float = 1.0, max_samples: int = 1024,
                 device="cuda:0", views_per_batch: Optional[int] = None, concurrent_batches: int = 3,
                 balance: str = "dynamic"):
        assert 1 <= len(radiance_fields) <= 4 and len(radiance_fields) == len(estimators)
        self.fields, self.estimators = list(radiance_fields), list(estimators)
        self.width, self.height, self.focal = int(width), int(height), float(focal)
        self.opts = dict(near_plane=near_plane, render_step_size=render_step_size, cone_angle=cone_angle,
                         alpha_thre=alpha_thre, max_samples=max_samples)
        self.device = torch.device(device)
        self.n_sem = self.fields[0].num_semantic_classes
        assert balance in ("lpt", "dynamic", "contiguous")
        self.balance = balance
        self.after_render = None  # optional callable(renderer), invoked once per finished render (measurement)
        self._probe = None
        self.last_probe_ms = 0.0  # wall time of the last cost probe (it ends with the host read of the costs)
        self._calls = 0
        PredictiveInformationScorer._instances += 1
        self._uid = PredictiveInformationScorer._instances  # the same on every rank (scorers are created in lock step)
        # rounded-linspace subsample of the full image (habitat_to_data.py:462-467)
        h, w = int(height * scale), int(width * scale)
        self.rays_per_view = h * w
        if self.rays_per_view == width * height:
            self.keep_idx = None
        else:
            idx = np.round(np.linspace(0, width * height - 1, self.rays_per_view)).astype(np.int32)
            self.keep_idx = torch.from_numpy(idx).to(self.device)
        # views per renderer pass: about 5 M rays (64 views of 320x240; 1200 views of 64x64) unless given
        self.views_per_batch = int(views_per_batch) if views_per_batch else max(1, 4915200 // self.rays_per_view)
        # Renders in flight: `concurrent_batches` view batches x the ensemble members, each with its own
        # working set and stream.  The marching loops are independent, so the narrow, launch-bound tail
        # iterations of one render overlap the wide early iterations of the others.
        self.concurrent_batches = max(1, int(concurrent_batches))
        self.renderers = [[FusedRenderer(self.device, self.n_sem) for _ in self.fields]
                          for _ in range(self.concurrent_batches)]
        self.renderer = self.renderers[0][0]
        self._streams = None
        self.interleave = True  # False: one render after the other on the current stream (measurement)
        # rays per view of the scheduler's cost probe: 1/768 of the view (100 of 320x240), at least 64
        self.probe_rays = int(os.environ.get("APNERF_PROBE_RAYS", str(max(64, self.rays_per_view // 768))))
        self.probe_min_samples = int(os.environ.get("APNERF_PROBE_MIN_SAMPLES", "64"))
        self.probe_iters = int(os.environ.get("APNERF_PROBE_ITERS", "4"))
        self.probe_tail_factor = float(os.environ.get("APNERF_PROBE_TAIL", "3"))
        # "dynamic": 1 = a rank ahead of the average drawn cost waits (``_PassQueue.may_draw``).  Off by default: on 8 GPUs
        # it took the device-timed step from 56.3 to 54.4 ms but the end-to-end call (ranks in lock step after every
        # call's host read) from 58.3 to 76 ms in the one run there was budget for (profiles/r02_scaling.md)
        self.throttle = int(os.environ.get("APNERF_DRAW_THROTTLE", "0"))
        self.shared_passes_per_rank = int(os.environ.get("APNERF_SHARED_PASSES", "8"))  # "dynamic": passes per rank in the shared counter
        self.min_batches = int(os.environ.get("APNERF_MIN_BATCHES", "3"))  # renderer passes per rank when view costs are known
        self.tail_priority = int(os.environ.get("APNERF_TAIL_PRIORITY", "0"))  # 1: tails of renders on high-priority streams (measured: no gain)
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
                     sums: Optional[torch.Tensor] = None, process_group=None) -> torch.Tensor:
        """c2w [n_views, 3, 4] f32 and view_traj [n_views] i32 ON THE DEVICE, the WHOLE batch on every rank -> float64
        [n_traj, 4] sums of the per-pixel (rgb, depth, sem, occ) predictive-information terms over the views THIS rank
        rendered (all of them without torch.distributed; ``schedule`` decides which and in which order: heaviest
        first).
        Everything is enqueued on CUDA streams; nothing is read back except the renderers' non-blocking look at their
        live-ray counters and the cost proxy's [n_views] counts."""
        n_views = c2w.shape[0]
        if sums is None:
            sums = torch.zeros((n_traj, 4), device=self.device, dtype=torch.float64)
        queue = self.schedule(c2w, process_group)
        self.views_rendered = 0
        self._buffers(min(n_views, queue.max_pass))
        # every renderer that can get a pass is sized for the LARGEST pass of the plan now: with passes drawn from the
        # shared counter a slot may meet its largest pass only steps later, and growing then stalls the device
        single_level = all(e.binaries.shape[0] == 1 for e in self.estimators)
        for slot in self.renderers[:max(1, min(self.concurrent_batches, queue.n_passes))]:
            for r in slot:
                if single_level:
                    r.reserve(min(n_views, queue.max_pass) * self.rays_per_view, min(n_views, queue.max_pass),
                              self.opts["cone_angle"])
        E, K = len(self.fields), self.concurrent_batches
        with torch.cuda.device(self.device):
            if self._streams is None:
                self._streams = [[torch.cuda.Stream(device=self.device) for _ in range(E)] for _ in range(K)]
                # the tail of a render continues on a HIGH-PRIORITY stream (see the rolling pipeline below)
                self._tail_streams = [[torch.cuda.Stream(device=self.device, priority=-1) for _ in range(E)]
                                      for _ in range(K)]
            main = torch.cuda.current_stream()
            free_slots = list(range(K))
            active = []  # batches in flight: dict(slot, views, nr, states, gens = [[stream, generator, iterations]])

            def start(views, slot):
                nv = int(views.shape[0])
                self.views_rendered += nv
                nr = nv * self.rays_per_view
                rays_o, rays_d = self._rays[slot][0][:nr], self._rays[slot][1][:nr]
                call("apnerf_generate_rays", nv, c2w.index_select(0, views).contiguous(), self.width, self.height,
                     self.focal, self.rays_per_view, self.keep_idx, rays_o, rays_d)
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
                                                   **self.opts), 0, r, self._tail_streams[slot][m]])
                return dict(slot=slot, views=views, nr=nr, states=states, gens=gens)

            def finish(job):
                for stream, *_ in job["gens"]:
                    done = torch.cuda.Event()
                    done.record(stream)
                    main.wait_event(done)
                if self.after_render is not None and self.interleave:
                    for r in self.renderers[job["slot"]]:
                        self.after_render(r)
                states = job["states"] + [None] * (4 - len(job["states"]))
                call("apnerf_score_views", E, states[0], states[1], states[2], states[3], job["nr"], self.rays_per_view,
                     self.n_sem, view_traj.index_select(0, job["views"]).contiguous(), n_traj, sums)
                free_slots.append(job["slot"])

            # Rolling pipeline.  The first dozen marching iterations of a batch are wide, throughput-bound launches; the
            # long tail (a few views whose rays cross much transparent occupied space keep marching 4 samples at a time,
            # up to 256 iterations) is a train of small latency-bound launches that leave most of the GPU idle.  A new
            # batch is therefore started, on its own streams, as soon as every batch in flight has left its head phase:
            # the next head's big kernels fill the SMs the tails do not use.
            while queue.has_more() or active:
                head_done = all(g[2] >= self.stagger_iters or g[1] is None for job in active for g in job["gens"])
                if queue.has_more() and free_slots and (not active or head_done) and queue.may_draw(bool(active)):
                    views = queue.next()
                    if views is not None:
                        active.append(start(views, free_slots.pop(0)))
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
                        if g[1] is not None and g[2] == self.stagger_iters and self.tail_priority:
                            # The render leaves its head phase: its remaining ~100 iterations are a serial chain of small
                            # launches -- the critical path of the pass.  They continue on a high-priority stream, so each
                            # of them waits at most for the kernel that is running, not behind the wide kernels the other
                            # passes have queued (measured: a rank's 110-iteration view took 58 ms beside other passes and
                            # 30 ms alone).
                            hop = torch.cuda.Event()
                            hop.record(g[0])
                            g[4].wait_event(hop)
                            g[0] = g[4]
                            with torch.cuda.stream(g[0]):
                                g[3].use_current_stream()
                    if not running:
                        active.remove(job)
                        finish(job)
        return sums

    def plan_batches(self, order: np.ndarray, cost: Optional[np.ndarray], sharing_ranks: int = 1):
        """Cut the views this rank may render (``order``) into the renderer passes of the rolling pipeline.
        Without costs (one GPU, or contiguous slices): even passes of at most ``views_per_batch`` views (72 views ->
        36 + 36, not 64 + 8), pass b = every n-th view.  With the probe's costs (``order`` is then heaviest first): passes
        of about EQUAL COST, consecutive in that order, at least ``min_batches`` of them -- so a view that is far heavier
        than the rest (a camera inside an occupied, transparent region: 20x the median cost, a serial chain of ~110
        four-sample iterations) gets a pass of its own that starts FIRST, and its long launch-bound tail overlaps the
        wide kernels of the passes behind it instead of following them (in one pass all views advance in lock step and
        the tail of the heaviest is exposed at the end).  ``sharing_ranks`` > 1: the passes are drawn from a counter
        shared by that many ranks ("dynamic"), so there are about ``shared_passes_per_rank`` per rank."""
        n = len(order)
        if n == 0:
            return []
        if cost is None:
            n_batches = max(1, -(-n // self.views_per_batch))
            return [order[b::n_batches] for b in range(n_batches)]
        want = max(self.min_batches, -(-n // self.views_per_batch),
                   self.shared_passes_per_rank * sharing_ranks if sharing_ranks > 1 else 1)
        want = min(want, n)
        c = np.maximum(np.asarray(cost, dtype=np.float64)[order], 0.0) + 1e-9
        batches, start, left = [], 0, float(c.sum())
        for b in range(want):
            if start >= n:
                break
            target = left / (want - b)
            end, acc = start, 0.0
            while end < n and end - start < self.views_per_batch and (end == start or acc + 0.5 * c[end] <= target):
                acc += c[end]
                end += 1
            if b == want - 1:  # the last pass takes what is left (further passes follow if it exceeds views_per_batch)
                end = min(n, start + self.views_per_batch)
                acc = float(c[start:end].sum())
            batches.append(order[start:end])
            left -= acc
            start = end
        while start < n:
            batches.append(order[start:start + self.views_per_batch])
            start += self.views_per_batch
        return batches

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

    # ---- scheduling: heavy views first, batches handed out dynamically --------------------------------------------
    # ---- scheduling over ranks: a cost probe, then cost-aware passes ------------------------------------------------------
    @torch.no_grad()
    def view_cost_proxy(self, c2w: torch.Tensor) -> np.ndarray:
        """Every view's cost ([n_views] float64, host) from a PROBE RENDER: `probe_rays` rays per view (a rounded-linspace
        subsample, the reference's own rule) are rendered through every ensemble member with the real renderer -- real
        occupancy grid, real density field, so rays terminate where the full-resolution ones will -- in a few COARSE
        marching iterations (`probe_min_samples` = 64 samples per ray and iteration instead of the reference's 4, stopped
        after `probe_iters` = 4 of them), and the sample rows each view sends through the field are counted on the device
        (`call_rows` of the schedule kernel); a ray still alive at the cut-off is counted as needing `probe_tail_factor`
        = 3 times as many samples again (such rays sit in transparent occupied space; the heaviest bench view needs
        880 samples per pixel over both members and the probe then says 1024).  Field
        evaluations are what a view costs (the field kernel is half of the step; marcher and compositor scale with the
        same count) and they vary 40x between poses: a camera inside an occupied but transparent region marches ~440
        samples per ray from the near plane on, 20x the median view, as a serial chain of ~110 four-sample iterations.
        Finding THOSE views is what the probe is for (they must start first); the split itself is self-balancing
        (``schedule``).  The un-truncated probe (min 16 samples, to the end) correlates 0.999 with the full-resolution
        count (profiles/r02_scaling.md) but is a serial chain of ~60 small launches = 6 ms on every rank; this one is four
        iterations.  The occupancy-only march count used before could not see transparency at all: the split it produced
        was WORSE balanced than contiguous slices.  Deterministic (integer counts of deterministic kernels), so every rank
        computes the same costs without talking."""
        n_views = c2w.shape[0]
        if any(e.binaries.shape[0] != 1 for e in self.estimators) or n_views == 0:
            return np.zeros(n_views)
        if self._probe is None:
            k = int(min(self.rays_per_view, max(16, self.probe_rays)))
            idx = np.round(np.linspace(0, self.rays_per_view - 1, k)).astype(np.int64)
            if self.keep_idx is not None:
                idx = self.keep_idx.cpu().numpy().astype(np.int64)[idx]
            self._probe = dict(k=k, keep=torch.from_numpy(idx.astype(np.int32)).to(self.device),
                               renderers=[FusedRenderer(self.device, self.n_sem) for _ in self.fields],
                               streams=[torch.cuda.Stream(device=self.device) for _ in self.fields])
        k, keep = self._probe["k"], self._probe["keep"]
        n_rays = n_views * k
        dev = self.device
        rays_o, rays_d = torch.empty((n_rays, 3), device=dev), torch.empty((n_rays, 3), device=dev)
        rows = torch.zeros((len(self.fields), n_views), device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            call("apnerf_generate_rays", n_views, c2w.contiguous(), self.width, self.height, self.focal, k, keep, rays_o,
                 rays_d)
            main = torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(main)
            gens = []
            for m, (f, e) in enumerate(zip(self.fields, self.estimators)):
                st = self._probe["streams"][m]
                st.wait_event(ready)
                gens.append([st, self._probe["renderers"][m].render_iter(
                    f, e, rays_o, rays_d, k, probabilistic=False, call_rows=rows[m], min_samples=self.probe_min_samples,
                    poll_every=0, **self.opts)])
            for _ in range(self.probe_iters):  # the members' probes advance in lock step on their own streams
                for g in gens:
                    if g[1] is not None:
                        with torch.cuda.stream(g[0]):
                            if next(g[1], None) is None:
                                g[1] = None
            alive = []
            for m, (st, g) in enumerate(gens):  # abandoned mid-render: n_alive_acc still holds the live rays per view
                with torch.cuda.stream(st):
                    alive.append(self._probe["renderers"][m].n_alive_acc[:n_views].clone() if g is not None
                                 else torch.zeros(n_views, device=dev, dtype=torch.int32))
                done = torch.cuda.Event()
                done.record(st)
                main.wait_event(done)
            cost = rows.sum(0).double() + torch.stack(alive).sum(0).double() * (
                self.probe_tail_factor * self.probe_iters * self.probe_min_samples)
            return cost.cpu().numpy()

    def schedule(self, c2w: torch.Tensor, process_group=None):
        """-> the queue of renderer passes of this rank.  One rank: all views, even passes.  "contiguous": this rank's
        slice, even passes.  Otherwise the cost probe runs first (``view_cost_proxy``; ~1 ms, the same numbers on every
        rank) and
          "dynamic" (default): ALL views are cut into equal-cost passes, heaviest first (`shared_passes_per_rank` per rank),
             and the ranks draw them from a counter in the process group's store whenever a renderer slot frees up: the
             heavy views start first on different ranks, and what the cost model cannot see (rows in the long launch-bound
             tail of a heavy view cost ~1.3x the rows of a wide iteration) is absorbed by who draws next.  A rank whose
             drawn cost is AHEAD of the average over the ranks (known from the counter: the passes are the same list on
             every rank) does not draw while it still has a pass in flight (``_PassQueue.may_draw``): kernels of
             concurrent passes do not overlap on the SMs (the persistent field kernel takes every SM), so a rank that
             keeps feeding its free slots while it renders a 20x view finishes that view -- and the step -- late
             (measured on 8 GPUs: every balanced policy ended at the 58 ms of the rank holding that view);
          "lpt": views dealt out longest-processing-time first to the least loaded rank (the same deterministic split on
             every rank, no communication at all), each rank's share cut into equal-cost passes, heaviest first."""
        import torch.distributed as dist

        n_views = c2w.shape[0]
        use_dist = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        rank = dist.get_rank(process_group) if use_dist else 0
        world = dist.get_world_size(process_group) if use_dist else 1
        self._calls += 1
        queue = _PassQueue(self.device)
        multi_level = any(e.binaries.shape[0] != 1 for e in self.estimators)
        if world == 1:  # nothing to balance; the order of the passes does not matter on one GPU (measured)
            queue.add_local(self.plan_batches(np.arange(n_views), None))
            return queue
        if self.balance == "contiguous" or n_views < 2 * world or multi_level:
            lo, hi = shard_range(n_views, rank, world)
            queue.add_local(self.plan_batches(np.arange(lo, hi), None))
            return queue
        t0 = time.perf_counter()
        cost = self.view_cost_proxy(c2w)
        self.last_cost, self.last_probe_ms = cost, 1e3 * (time.perf_counter() - t0)
        if self.balance == "dynamic":
            passes = self.plan_batches(np.argsort(-cost, kind="stable"), cost, world)
            store = dist.distributed_c10d._get_default_store()
            queue.set_shared(passes, _Tickets(len(passes), store, f"apnerf/tickets/{self._uid}/{self._calls}"),
                             [float(cost[b].sum()) for b in passes] if self.throttle else None)
            return queue
        mine = lpt_assign(cost, world)[rank]
        queue.add_local(self.plan_batches(mine[np.argsort(-cost[mine], kind="stable")], cost))
        return queue

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
        c2w_host = torch.from_numpy(poses_to_c2w(poses)).pin_memory()
        vt_host = torch.from_numpy(np.ascontiguousarray(np.asarray(view_traj), dtype=np.int32)).pin_memory()
        c2w = c2w_host.to(self.device, non_blocking=True)
        vt = vt_host.to(self.device, non_blocking=True)
        sums = torch.zeros((n_traj, 4), device=self.device, dtype=torch.float64)
        if poses.shape[0]:
            self.partial_sums(c2w, vt, n_traj, sums, process_group)
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


def lpt_assign(cost: np.ndarray, world: int, initial=None):
    """Longest-processing-time-first assignment of units with the given costs to `world` bins (that already hold the
    loads `initial`); returns one ascending index array per bin.  Deterministic (stable sort, ties to the lower bin), so
    every rank computes the same split."""
    order = np.argsort(-np.asarray(cost, dtype=np.float64), kind="stable")
    load = np.zeros(world) if initial is None else np.asarray(initial, dtype=np.float64).copy()
    bins = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += max(float(cost[i]), 0.0) + 1e-9
    return [np.asarray(sorted(b), dtype=np.int64) for b in bins]


class _PassQueue:
    """The renderer passes a rank still has to start: its own list, then (once the cost probe has planned them) the
    passes all ranks share, drawn from the ticket counter."""

    def __init__(self, device):
        self.device, self.local, self.shared, self.tickets, self.max_pass, self.n_passes = device, [], None, None, 1, 0

    def _dev(self, passes):
        self.max_pass = max([self.max_pass] + [len(b) for b in passes])
        self.n_passes += sum(1 for b in passes if len(b))
        return [torch.from_numpy(np.ascontiguousarray(b, dtype=np.int64)).to(self.device) for b in passes if len(b)]

    def add_local(self, passes):
        self.local += self._dev(passes)

    def set_shared(self, passes, tickets, pass_costs=None):
        """pass_costs (optional, one per pass, the same on every rank): enables the draw throttle (``may_draw``)."""
        keep = [i for i, b in enumerate(passes) if len(b)]
        self.shared, self.tickets = self._dev(passes), tickets
        self.my_load, self._turn, self._blocked, self._heavy = 0.0, 0, False, False
        self.cum = None
        if pass_costs is not None and len(keep):
            c = np.asarray(pass_costs, dtype=np.float64)[keep]
            mean = float(c.sum()) / len(c)
            # a pass far above the mean is ONE heavy view (the planner cuts equal-cost passes): its rows sit in a long
            # launch-bound tail and cost ~1.3x the rows of wide iterations (profiles/r02_scaling.md)
            c = np.where(c > 2.0 * mean, 1.3 * c, c)
            self.pass_costs, self.cum, self.slack, self.heavy_above = c, np.concatenate([[0.0], np.cumsum(c)]), mean, 2.0 * mean

    def may_draw(self, has_active: bool, peek_every: int = 6) -> bool:
        """Draw throttle of the shared queue: a rank that already has a pass in flight draws another one only while the
        cost it has drawn so far is at most one average pass above the average over the ranks -- and, once it has drawn a
        HEAVY pass (one view far above the mean), only while it is not above that average at all: the heavy view is the
        step's critical path and every pass rendered beside it delays it by its whole kernel time.  The average is known
        without talking to anybody but the counter: passes [0, b) have been drawn, and every rank holds the same list
        of pass costs.  A rank with nothing in flight always draws, so no rank ever idles and nothing can dead-lock.
        The counter is looked at (one store round trip) at most every `peek_every` calls."""
        if self.local or self.shared is None or self.cum is None or not has_active:
            return True
        self._turn += 1
        if self._blocked and self._turn % peek_every:
            return False
        b = min(self.tickets.peek(), len(self.shared))
        self._blocked = self.my_load > self.cum[b] / self.tickets.world + (0.0 if self._heavy else self.slack)
        return not self._blocked

    def has_more(self) -> bool:
        return bool(self.local) or self.shared is not None

    def next(self):
        """The next pass (device tensor of view indices) or None."""
        if self.local:
            return self.local.pop(0)
        if self.shared is not None:
            b = self.tickets.take(1)
            if b < len(self.shared):
                if self.cum is not None:
                    self.my_load += float(self.pass_costs[b])
                    self._heavy = self._heavy or self.pass_costs[b] > self.heavy_above
                return self.shared[b]
            self.shared = None
        return None


class _Tickets:
    """The counter the ranks draw view batches from.  One process: a local integer.  Several ranks: a key of the
    process group's store (``store.add`` is atomic; one TCP round trip of ~0.1 ms per draw, a handful of draws per rank
    and call) -- no collective, so a rank never waits for a slower one until the final all-reduce of the sums."""

    def __init__(self, n: int, store, key: str):
        self.n, self.store, self.key, self.next = n, store, key, 0
        self.world = 1
        if store is not None:
            import torch.distributed as dist

            self.world = dist.get_world_size()

    def peek(self) -> int:
        """How many positions have been handed out so far (one store round trip with several ranks)."""
        return self.next if self.store is None else int(self.store.add(self.key, 0))

    def take(self, b: int) -> int:
        """Reserve the next ``b`` positions; returns the first one (>= n when nothing is left)."""
        if self.store is None:
            t, self.next = self.next, self.next + b
            return t
        return int(self.store.add(self.key, b)) - b


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
                                             alpha_thre=alpha_thre, scale=scale, device=device)
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
