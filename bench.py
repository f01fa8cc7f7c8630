#!/usr/bin/env python
"""Benchmark of the render + score hot path (BASELINE.json metric: rays/s rendered + scored).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload score|train|round]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--workload score (default; BASELINE.json configs[2]): the planner candidate-view batch -- 256 synthetic poses
    (positions U(aabb shrunk by 1 m), y = 1.5, yaw U[0, 2 pi), seed 3; 8 trajectories x 32 views)
    x 320x240 rays x 2 ensemble members through a 128^3 occupancy grid and a 16-level hash-grid NeRF with 29 semantic
    classes, "trained-like" random weights (synthetic.py), rendered with the reference's
    test-mode schedule (max_samples 1024) and reduced to predictive information per trajectory.  One "ray" = one
    pixel of one view through one member.  A step = the whole 256-pose batch; at N > 1 the SAME 256 poses are
    sharded over the ranks (strong scaling; `--scaling weak --views-per-gpu V` gives V poses per rank instead).
--workload train (configs[3]): one data-parallel NeRF training step per member-model, 8192 rays per batch per GPU.
--workload round (configs[4]): score 20 trajectories x 40 views x 2 members, argmax, then retrain both members.

value : device-resident throughput (poses already in HBM, scores left in HBM).
e2e   : the same through PredictiveInformationScorer.score_views with HOST pose arrays:
        pose -> matrix on the host, pinned H2D, render + score, all-reduce, D2H of the scores.
The per-step working set (ray state + per-iteration sample buffers, several GB) is far larger than the 126 MB L2,
so no explicit L2 flush is needed between iterations (config.l2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, HFOV_FOCAL = 320, 240, 160.0
OPTS = dict(near_plane=0.1, render_step_size=1e-3, cone_angle=0.004, alpha_thre=0.01)
METRIC = "rays/s rendered+scored (pred-info)"
N_SEM = 29
FIELD_SEEDS = (2, 12)
VIEWS_PER_TRAJ = 32
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_field_kernel_ncu.json")  # written from the committed ncu capture


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1643.1)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1643.1, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index
        self.lo = self.hi = None

    def mark_begin(self):
        self.lo = len(self.lines)

    def mark_end(self):
        self.hi = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        lines = self.lines[self.lo:self.hi] if self.lo is not None else self.lines
        window = "timed region"
        if len(lines) < 3:  # region shorter than a few sampling periods: use everything since warm-up began
            lines, window = self.lines, "warm-up + timed region + e2e (timed region < 3 samples)"
        for l in lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def _dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def _scene(apnerf, synthetic, dev, density_gain, train_mode=False):
    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
    est.binaries = synthetic.make_occupancy(128, seed=1)
    est.occs = est.binaries.flatten().float() * 0.5  # a consistent EMA state for the training workloads
    est = est.to(dev)
    fields = []
    for s in FIELD_SEEDS:
        f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=N_SEM)
        fields.append(synthetic.init_trained_like(f, seed=s, density_gain=density_gain).to(dev))
    if not train_mode:
        est.eval()
        [f.eval() for f in fields]
    return est, fields


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same path (bounded sample: whole 320x240 views)
# ------------------------------------------------------------------------------------------
class CpuPath:
    """oracle/ (C + numpy restatement of the reference path) on host cores: whole 320x240 views of the bench's own
    pose set through both ensemble members (the per-call schedule n = R // n_alive depends on R, so the CPU arm
    marches the same 76 800-ray calls as the GPU arm), then the float64 scoring."""

    def __init__(self, density_gain):
        import torch
        from oracle import oracle as O
        import apnerf
        from apnerf import synthetic

        O.build()
        self.O, self.synthetic = O, synthetic
        torch.set_num_threads(os.cpu_count() or 1)
        self.occ = synthetic.make_occupancy(128, seed=1).numpy()
        self.aabbs = np.asarray([synthetic.ROI_AABB], np.float32)
        self.fns = []
        for s in FIELD_SEEDS:
            f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=N_SEM)
            synthetic.init_trained_like(f, seed=s, density_gain=density_gain)
            fp = O.FieldParams(f.mlp_base.params.detach().numpy(), f.mlp_head.params.detach().numpy(),
                               f.mlp_sem.params.detach().numpy(), num_semantic_classes=N_SEM)
            aabb = np.asarray(synthetic.ROI_AABB, np.float32)
            self.fns.append(lambda p, d, fp=fp, aabb=aabb: O.field_forward(p, d, aabb, fp))
        self.poses = synthetic.make_poses(256, seed=3)

    def view(self, v):
        """Render + score view v (both members): returns (rays, evaluated+composited samples, seconds)."""
        O = self.O
        t0 = time.perf_counter()
        o, d = O.generate_image_rays(self.synthetic.pose_to_matrix(self.poses[v % 256]).astype(np.float32), W, H, HFOV_FOCAL)
        outs, n_samples = [], 0
        for fn in self.fns:
            r = O.render_probablistic_image_with_occgrid_test(1024, fn, self.occ, self.aabbs, o, d, N_SEM, **OPTS)
            outs.append(r)
            n_samples += r[6]
        stack = lambda k: np.stack([r[k] for r in outs])[:, None]
        O.predictive_information(stack(1), stack(4)[..., 0], stack(2)[..., 0], stack(5))
        return W * H * len(self.fns), n_samples, time.perf_counter() - t0


def cpu_baseline(density_gain, n_views=1):
    cp = CpuPath(density_gain)
    rays = samples = 0
    secs = 0.0
    for v in range(n_views):
        r, s, dt = cp.view(v)
        rays, samples, secs = rays + r, samples + s, secs + dt
    return dict(value=rays / secs, unit="rays/s", cores=int(cp.O.N_THREADS), kind="port",
                sample=f"{n_views} whole 320x240 view(s) of the bench's pose set x 2 members = {rays} rays, {samples} "
                       f"composited samples, {secs:.1f} s; oracle/ C+numpy port of the reference path on all host "
                       f"cores (the reference's own CUDA/tcnn path has no CPU implementation)")


def _score_static_config(args, world):
    """The part of `config` that names the workload (identical for the GPU arm and the reference arm)."""
    V_total = args.views if args.scaling == "strong" else args.views_per_gpu * world
    return {"workload": f"planner candidate-view batch (BASELINE.json configs[2]): {V_total} poses x {W}x{H} rays "
                        f"x 2 ensemble members, 128^3 occ grid, 16-level hash NeRF, sem-num 29, max_samples "
                        f"1024, render+score (pred-info); {'the same poses sharded' if args.scaling == 'strong' else 'poses per rank fixed'} "
                        f"over {world} rank(s)",
            "views_total": V_total, "rays_per_step": V_total * W * H * len(FIELD_SEEDS), "ensemble": len(FIELD_SEEDS),
            "n_trajectories": max(1, V_total // VIEWS_PER_TRAJ),
            "poses": "SURVEY 8(d)-3: x,z U(aabb shrunk by 1 m), y 1.5, yaw U[0,2pi), seed 3",
            "init": f"hash features U(-1,1), Xavier MLPs, density row x{args.density_gain} (BASELINE.md section 6), "
                    f"field seeds {FIELD_SEEDS}",
            "density_gain": args.density_gain,
            "l2": "per-step working set (> 10 GB at 256 poses) exceeds the 126 MB L2; no flush needed"}


def run_reference(args):
    rank, _, _ = _dist_env()
    if rank != 0:
        return
    cp = CpuPath(args.density_gain)
    times, rays, comp = [], 0, 0
    for i in range(args.warmup + args.steps):
        r, ns, dt = cp.view(i)
        if i >= args.warmup:
            times.append(dt)
            rays += r
            comp += ns
    total = sum(times)
    v = rays / total
    cb = dict(value=v, unit="rays/s", cores=int(cp.O.N_THREADS), kind="port",
              sample=f"each step: one whole 320x240 view of the bench's pose set x 2 members (153 600 rays) + scoring; "
                     f"views {args.warmup}..{args.warmup + args.steps - 1}, {comp / max(1, rays):.1f} composited samples "
                     f"per ray; {args.steps} timed steps, {total:.1f} s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "fp32 (fp16-rounded MLP operands) / f64 scoring",
        "data": "synthetic",
        "config": dict(_score_static_config(args, max(1, args.gpus)),
                       sample="bounded CPU sample per step: 1 whole 320x240 view x 2 members of the workload's pose set",
                       sample_rays_per_step=rays // max(1, args.steps)),
        "cpu_baseline": cb, "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------
# GPU arm: render + score
# ------------------------------------------------------------------------------------------
def run_score(args):
    import torch
    import torch.distributed as dist

    import apnerf
    from apnerf import _lib, synthetic

    rank, world, local = _dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R = W * H
    V_total = args.views if args.scaling == "strong" else args.views_per_gpu * world
    est, fields = _scene(apnerf, synthetic, dev, args.density_gain)
    n_traj = max(1, V_total // VIEWS_PER_TRAJ)
    poses_all = synthetic.make_poses(V_total, seed=3)
    view_traj_all = np.minimum(np.arange(V_total) // VIEWS_PER_TRAJ, n_traj - 1).astype(np.int32)
    scorer = apnerf.PredictiveInformationScorer(fields, [est, est], W, H, HFOV_FOCAL, device=dev,
                                                views_per_batch=args.views_per_batch or None,
                                                concurrent_batches=args.concurrent_batches,
                                                balance=args.balance, **OPTS)
    c2w = torch.from_numpy(apnerf.scoring.poses_to_c2w(poses_all)).to(dev)  # the whole batch on every rank (12 KB)
    vt = torch.from_numpy(view_traj_all).to(dev)
    sums = torch.zeros((n_traj, 4), device=dev, dtype=torch.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    own = []  # (start, end-before-the-all-reduce) events of every step: this rank's OWN render + score time

    def step_device():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sums.zero_()
        scorer.partial_sums(c2w, vt, n_traj, sums)
        b.record()
        own.append((a, b))
        if world > 1:
            dist.all_reduce(sums)

    def step_e2e():
        return scorer.score_views(poses_all, view_traj_all, n_traj)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler.mark_begin()
    _lib.LAUNCHES.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    apnerf.FusedRenderer.host_blocked_s = 0.0
    own.clear()
    h0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    host_s = time.perf_counter() - h0  # wall time of the enqueuing thread inside the timed steps (no sync in there)
    host_blocked_s = apnerf.FusedRenderer.host_blocked_s  # ... of which waiting for the device (the renderers' throttle)
    e1.record()
    barrier()
    sampler.mark_end()
    launches = _lib.kernel_launches()
    ms = e0.elapsed_time(e1)
    own_ms = sum(a.elapsed_time(b) for a, b in own)
    views_mine = scorer.views_rendered  # of the last step (dynamic: varies a little from step to step)
    # end to end through the public API (host poses in, host scores out)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        terms = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms, e2e_s * 1e3, 1e3 * (host_s - host_blocked_s), own_ms], device=dev, dtype=torch.float64)
    all_ms = [torch.zeros_like(tms) for _ in range(world)]
    if world > 1:
        dist.all_gather(all_ms, tms)
    else:
        all_ms = [tms]
    per_rank_ms = [float(t[3]) / args.steps for t in all_ms]  # each rank's own work, without the wait in the all-reduce
    per_rank_host_busy = [float(t[2]) / args.steps for t in all_ms]
    ms, e2e_ms = max(float(t[0]) for t in all_ms), max(float(t[1]) for t in all_ms)

    # ---- roofline of the dominant kernel (field_forward_kernel), measured live on every rank ----
    roof = field_kernel_roofline(torch, scorer, c2w, vt, n_traj)
    rows = torch.tensor([float(roof["samples_per_step"])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rows)
    total_rows = float(rows[0])
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.density_gain)
    rays_per_step = V_total * R * len(fields)
    if rank == 0:
        out = {
            "metric": METRIC, "value": rays_per_step * args.steps / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f16 MLP operands / f32 accumulate+compositing / f64 scoring", "data": "synthetic",
            "config": dict(_score_static_config(args, world), **{
                       "views_per_gpu": views_mine, "views_per_batch": scorer.views_per_batch,
                       "shard": {"lpt": "static split balanced by the cost of a low-resolution probe render (LPT), heaviest views first",
                                 "dynamic": "equal-cost passes drawn heaviest-first from a counter shared by the ranks",
                                 "contiguous": "contiguous slices"}[scorer.balance] if world > 1 else "one rank: all views",
                       "schedule_probe_ms": scorer.last_probe_ms,
                       "mean_samples_per_ray": total_rows / rays_per_step,
                       "parallelism": f"views sharded over {world} rank(s), one all-reduce of [n_traj,4] f64"}),
            "samples_per_s": total_rows * args.steps / (ms * 1e-3), "samples_per_step": total_rows,
            "per_rank_ms_per_step": {"min": min(per_rank_ms), "max": max(per_rank_ms)},
            "host_busy_ms_per_step": {"min": min(per_rank_host_busy), "max": max(per_rank_host_busy),
                                      "what": "enqueuing thread's wall time inside the timed steps minus its waits for the "
                                              "device; close to ms_per_step = host-bound"},
            "clocks": clocks,
            "e2e": {"value": rays_per_step * args.steps / (e2e_ms * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": int(V_total * (12 * 4 + 4)), "d2h_bytes_per_step": int(n_traj * 4 * 8)},
            "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "scores_sample": np.round(terms[0], 6).tolist(),
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def field_kernel_roofline(torch, scorer, c2w, vt, n_traj):
    """One instrumented pass: CUDA events around every field_forward launch (on the launching stream) + the rows the
    renderers evaluated -> achieved algorithmic GB/s of the hash-grid gather."""
    from apnerf import _lib

    hbm_peak, tensor_peak, peak_kind = _peaks()
    evs = []

    def hook(pc):
        if not pc.name.startswith("apnerf_field_forward"):
            return pc.invoke()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        pc.invoke()
        a1.record()
        evs.append((a0, a1))

    counted = []
    orig_hook = scorer.after_render
    scorer.after_render = lambda r: counted.append(r.counters[8:9].clone())
    _lib.CALL_HOOK = hook
    scorer.interleave = False  # time the members' kernels without cross-stream contention
    try:
        sums = torch.zeros((n_traj, 4), device=c2w.device, dtype=torch.float64)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        scorer.partial_sums(c2w, vt, n_traj, sums)
        t1.record()
        torch.cuda.synchronize()
    finally:
        _lib.CALL_HOOK = None
        scorer.interleave = True
        scorer.after_render = orig_hook
    times = [a.elapsed_time(b) for a, b in evs]
    k_ms = sum(times)
    rows = int(torch.cat(counted).sum().item()) if counted else 0
    n_launch = sum(1 for t in times if t > 0.02)  # launches that found work (an empty launch is ~8 us)
    step_ms = t0.elapsed_time(t1)
    bytes_per_sample = 1024  # 16 levels x 8 corners x 4 features x 2 B
    achieved = rows * bytes_per_sample / (k_ms * 1e-3) / 1e9
    ncu = json.load(open(NCU_SUMMARY)) if os.path.exists(NCU_SUMMARY) else None
    out = {"bound": "hbm", "limiter": "l2-gather: the 48 MB fp16 table is L2-resident, so the gather is bounded by the SM's "
                                      "L1/L2 request path, not by HBM; `achieved` is HBM-equivalent algorithmic bytes",
           "kernel": "field_forward_kernel (hash-grid gather + fused tcgen05 MLPs)",
           "achieved": achieved, "peak": hbm_peak, "peak_kind": peak_kind + ", HBM copy GB/s",
           "unit": "GB/s", "frac": achieved / hbm_peak,
           "traffic": (ncu["dram_bytes_per_row"] * rows / max(1, n_launch)) if ncu else None,
           "traffic_source": (f"profiles/{os.path.basename(NCU_SUMMARY)}: dram__bytes_read.sum + dram__bytes_write.sum per "
                              f"row of the committed ncu --set full capture x rows per launch of this run") if ncu else None,
           "ncu": ncu,
           "algorithmic_bytes_per_launch": bytes_per_sample * rows / max(1, n_launch),
           "algorithmic_bytes_per_sample": bytes_per_sample, "samples_per_step": rows,
           "launches_with_work": n_launch, "avg_launch_ms": k_ms / max(1, n_launch),
           "kernel_share_of_step": k_ms / step_ms, "gsamples_per_s": rows / (k_ms * 1e-3) / 1e9,
           "mlp_tflops": rows * 81920 / (k_ms * 1e-3) / 1e12,
           "mlp_frac_of_tensor_peak": rows * 81920 / (k_ms * 1e-3) / 1e12 / tensor_peak}
    return out


# ------------------------------------------------------------------------------------------
# GPU arm: training step (configs[3]) and full round (configs[4])
# ------------------------------------------------------------------------------------------
def _train_setup(args, torch, apnerf, synthetic, dev, rank, world):
    est, fields = _scene(apnerf, synthetic, dev, 2.0, train_mode=True)
    ests = [est] + [apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1).to(dev) for _ in fields[1:]]
    for e in ests[1:]:
        e.load_state_dict(est.state_dict())
    # pipeline.py:173-178; fused=True: the same update as one multi-tensor kernel (a PyTorch library op: A18 stays in PyTorch)
    opts = [torch.optim.Adam(f.parameters(), lr=1e-3, eps=1e-15, fused=True) for f in fields]
    data = synthetic.TrainingSet(n_images=40, width=W, height=H, focal=HFOV_FOCAL, n_classes=N_SEM, seed=4 + rank, device=dev)
    return fields, ests, opts, data


def run_train(args):
    import torch
    import torch.distributed as dist

    import apnerf
    from apnerf import _lib, synthetic, training

    rank, world, local = _dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fields, ests, opts, data = _train_setup(args, torch, apnerf, synthetic, dev, rank, world)
    n_rays = args.rays_per_batch
    trainer = training.EnsembleTrainer(fields, ests, opts, **OPTS)

    def step(i):
        # one optimisation step of every ensemble member on its own fresh ray batch (pipeline.py:403-532)
        return trainer.step(lambda: data.fetch(n_rays), 1000 + i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler.mark_begin()
    _lib.LAUNCHES.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_samples = 0
    logs = None
    for i in range(args.steps):
        logs = step(args.warmup + i)
    e1.record()
    barrier()
    sampler.mark_end()
    launches = _lib.kernel_launches()
    n_samples = trainer.samples_seen
    ms = e0.elapsed_time(e1)
    # e2e: the batch comes from HOST memory every step (pinned image tensors -> H2D), loss read back
    host = {k: v.cpu().pin_memory() for k, v in (("images", data.images), ("depths", data.depths), ("semantics", data.semantics))}

    def step_e2e(i):
        return trainer.step(lambda: data.fetch_from_host(host, n_rays), 2000 + i, read_loss=True)

    step_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(1 + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])
    rays_per_step = n_rays * world * len(fields)
    if rank == 0:
        print(json.dumps({
            "metric": "training rays/s (fwd + bwd + Adam, per ensemble member-model)", "value": rays_per_step * args.steps / (ms * 1e-3),
            "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 MLP operands / f32 accumulate, f32 master weights + Adam", "data": "synthetic",
            "config": {"workload": f"NeRF training step (BASELINE.json configs[3]): {n_rays} rays per batch per GPU per "
                                   f"member, 2 ensemble members per step, 40 synthetic posed 320x240 rgb/depth/semantic "
                                   f"images, loss 10 smoothL1(rgb) + smoothL1(depth)/5 + CE(sem)/2, Adam(1e-3, eps 1e-15), "
                                   f"occupancy-grid update every 16 steps; data-parallel over {world} rank(s)",
                       "rays_per_batch_per_gpu": n_rays, "ensemble": len(fields),
                       "mean_samples_per_member_step": n_samples / max(1, trainer.steps_done),
                       "parallelism": f"dp{world}: rays split over ranks, gradient all-reduce (NCCL)",
                       "l2": "activations + table gradient (> 200 MB per step) exceed the L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": rays_per_step * args.steps / (e2e_ms * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": int(n_rays * len(fields) * (3 + 4 + 8 + 8)), "d2h_bytes_per_step": 4 * len(fields)},
            "gpu_launches": launches, "loss": logs,
        }))
    if world > 1:
        dist.destroy_process_group()


def run_round(args):
    """configs[4]: N trajectories x K views render + score (argmax), then retrain both members, one timed call."""
    import torch
    import torch.distributed as dist

    import apnerf
    from apnerf import _lib, synthetic, training

    rank, world, local = _dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fields, ests, opts, data = _train_setup(args, torch, apnerf, synthetic, dev, rank, world)
    n_traj, traj_len = 20, 80
    trajs = [synthetic.make_poses(traj_len, seed=100 + t) for t in range(n_traj)]
    scale = args.round_scale
    Wf, Hf = (640, 640) if scale < 1 else (W, H)  # reference: 640x640 subsampled to 64x64 (scale 0.1); or full 320x240
    scorer = apnerf.PredictiveInformationScorer(fields, ests, Wf, Hf, Wf / 2.0, device=dev, scale=scale,
                                                views_per_batch=args.views_per_batch or None, balance=args.balance, **OPTS)
    trainer = training.EnsembleTrainer(fields, ests, opts, **OPTS)

    def one_round(i):
        [f.eval() for f in fields]
        [e.eval() for e in ests]
        terms = scorer.score_trajectories(trajs)
        best = int(np.argmax(terms.sum(1)))  # pipeline.py:1085
        for s in range(args.train_steps):
            trainer.step(lambda: data.fetch(args.rays_per_batch), i * args.train_steps + s)
        return best, terms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        one_round(i)
    barrier()
    sampler.mark_begin()
    _lib.LAUNCHES.clear()
    t0 = time.perf_counter()
    for i in range(args.steps):
        best, terms = one_round(args.warmup + i)
    barrier()
    dt = time.perf_counter() - t0
    sampler.mark_end()
    launches = _lib.kernel_launches()
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    dt = float(tms[0])
    rays_scored = n_traj * 40 * scorer.rays_per_view * len(fields)
    if rank == 0:
        print(json.dumps({
            "metric": "full active-perception round: seconds per round (score + retrain)", "value": dt / args.steps,
            "unit": "s/round", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 MLP operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": f"full round (BASELINE.json configs[4]): score {n_traj} trajectories x 40 views x 2 members at "
                                   f"{scorer.rays_per_view} rays/view, argmax, then {args.train_steps} training steps x 2 members "
                                   f"({args.rays_per_batch} rays/batch/GPU); host poses in, best index out",
                       "rays_scored_per_round": rays_scored, "train_steps": args.train_steps},
            "clocks": clocks,
            "e2e": {"value": dt / args.steps, "unit": "s/round", "h2d_bytes_per_step": int(n_traj * 40 * 52),
                    "d2h_bytes_per_step": int(n_traj * 32)},
            "gpu_launches": launches, "best_trajectory": best, "scores": np.round(terms.sum(1), 5).tolist(),
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="score", choices=["score", "train", "round"])
    ap.add_argument("--views", type=int, default=256, help="total poses of the batch (strong scaling)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--views-per-gpu", type=int, default=256, help="poses per rank with --scaling weak")
    ap.add_argument("--views-per-batch", type=int, default=0, help="views per renderer pass (0: about 5 M rays)")
    ap.add_argument("--concurrent-batches", type=int, default=3,
                    help="view batches rendered concurrently per GPU (each x the ensemble members, own streams)")
    ap.add_argument("--balance", default="dynamic", choices=["lpt", "dynamic", "contiguous"],
                    help="multi-GPU view assignment: equal-cost passes drawn heaviest-first from a shared counter (default), "
                         "static split balanced by the probe's costs, or contiguous slices")
    ap.add_argument("--density-gain", type=float, default=6.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rays-per-batch", type=int, default=8192)
    ap.add_argument("--train-steps", type=int, default=2000, help="--workload round: training steps per round")
    ap.add_argument("--round-scale", type=float, default=0.1, help="--workload round: 0.1 = 64x64 of 640x640; 1 = 320x240")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    elif args.workload == "round":
        run_round(args)
    else:
        run_score(args)


if __name__ == "__main__":
    main()
