#!/usr/bin/env python
"""Benchmark of the render + score hot path (BASELINE.json metric: rays/s rendered + scored).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): the planner candidate-view batch of BASELINE.json configs[2] --
32 synthetic poses per GPU (256 at 8 GPUs) x 320x240 rays x 2 ensemble members through a
128^3 occupancy grid and a 16-level hash-grid NeRF with 29 semantic classes, random
"trained-like" weights (synthetic.py), rendered with the reference's test-mode schedule
(max_samples 1024) and reduced to predictive information per trajectory.  One "ray" = one pixel
of one view through one member.  A step = one such batch per GPU (weak scaling).

value : device-resident throughput (poses already in HBM, scores left in HBM).
e2e   : the same through PredictiveInformationScorer.score_views with HOST pose arrays:
        pose -> matrix on the host, pinned H2D, render + score, all-reduce, D2H of the scores.
The per-step working set (ray state 0.75 GB + sample buffers 1.4 GB) is far larger than the
126 MB L2, so no explicit L2 flush is needed between iterations (config.l2).
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
NCU_DRAM_BYTES_PER_SAMPLE = 104.4  # profiles/r01_field_kernel.md (dram__bytes_read.sum + write.sum per sample)
N_SEM = 29


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


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
        self.window = window
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
                "reasons": sorted(reasons), "samples": len(sm), "window": self.window}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same path (bounded sample)
# ------------------------------------------------------------------------------------------
def cpu_path_rays_per_s(min_seconds=10.0, rays_per_view=4096, max_views=8, seed_fields=(2, 12)):
    """Times oracle/ (C + numpy restatement of the reference path) on host cores: views of the
    same synthetic scene, subsampled to 64x64 rays by the reference's own rounded linspace
    (habitat_to_data.py:462-467), both ensemble members, then the float64 scoring."""
    import torch
    from oracle import oracle as O
    import apnerf
    from apnerf import synthetic

    O.build()
    torch.set_num_threads(os.cpu_count() or 1)
    occ = synthetic.make_occupancy(128, seed=1).numpy()
    aabbs = np.asarray([synthetic.ROI_AABB], np.float32)
    fns = []
    for s in seed_fields:
        f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=N_SEM)
        synthetic.init_trained_like(f, seed=s)
        fp = O.FieldParams(f.mlp_base.params.detach().numpy(), f.mlp_head.params.detach().numpy(),
                           f.mlp_sem.params.detach().numpy(), num_semantic_classes=N_SEM)
        aabb = np.asarray(synthetic.ROI_AABB, np.float32)
        fns.append(lambda p, d, fp=fp, aabb=aabb: O.field_forward(p, d, aabb, fp))
    poses = synthetic.make_poses(max_views, seed=3)
    keep = O.subsample_indices(W * H, rays_per_view)
    t0 = time.perf_counter()
    n_rays, n_samples, outs = 0, 0, [[], []]
    for v in range(max_views):
        o, d = O.generate_image_rays(synthetic.pose_to_matrix(poses[v]).astype(np.float32), W, H, HFOV_FOCAL)
        o, d = o[keep], d[keep]
        for m, fn in enumerate(fns):
            r = O.render_probablistic_image_with_occgrid_test(1024, fn, occ, aabbs, o, d, N_SEM, **OPTS)
            outs[m].append(r)
            n_rays += rays_per_view
            n_samples += r[6]
        if time.perf_counter() - t0 >= min_seconds:
            break
    nv = len(outs[0])
    stack = lambda k: np.stack([np.stack([outs[m][v][k] for v in range(nv)]) for m in range(2)])
    O.predictive_information(stack(1), stack(4)[..., 0], stack(2)[..., 0], stack(5))
    dt = time.perf_counter() - t0
    return dict(value=n_rays / dt, unit="rays/s", cores=int(O.N_THREADS), kind="port",
                sample=f"{nv} view(s) x {rays_per_view} rays (64x64 rounded-linspace subsample of 320x240) x 2 members, "
                       f"{n_samples} composited samples, {dt:.1f} s; oracle/ C+numpy port of the reference path "
                       f"(the reference's own CUDA/tcnn path has no CPU implementation)"), dt, n_rays


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        cb, dt, n = cpu_path_rays_per_s(min_seconds=0.0, max_views=1)
        if i >= args.warmup:
            times.append(dt)
            rays += n
    total = sum(times)
    v = rays / total
    cb["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (fp16-rounded MLP operands)", "data": "synthetic",
        "config": {"workload": "planner candidate-view batch (BASELINE.json configs[2]), bounded CPU sample per step: "
                               "1 view x 4096 rays x 2 members", "rays_per_step": rays // max(1, args.steps)},
        "cpu_baseline": cb, "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import apnerf
    from apnerf import _lib, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    V = args.views_per_gpu
    R = W * H
    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
    est.binaries = synthetic.make_occupancy(128, seed=1)
    est = est.to(dev).eval()
    fields = []
    for s in (2, 12):
        f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=N_SEM)
        fields.append(synthetic.init_trained_like(f, seed=s).to(dev).eval())
    scorer = apnerf.PredictiveInformationScorer(fields, [est, est], W, H, HFOV_FOCAL, device=dev,
                                                views_per_batch=max(1, V // args.concurrent_batches),
                                                concurrent_batches=args.concurrent_batches, **OPTS)
    n_traj = max(1, (V * world) // 32)
    poses_all = synthetic.make_poses(V * world, seed=3)
    view_traj_all = (np.arange(V * world) // 32).astype(np.int32) if V * world >= 32 else np.zeros(V * world, np.int32)
    lo, hi = apnerf.scoring.shard_range(V * world, rank, world)
    c2w = torch.from_numpy(apnerf.scoring.poses_to_c2w(poses_all[lo:hi])).to(dev)
    vt = torch.from_numpy(view_traj_all[lo:hi]).to(dev)
    sums = torch.zeros((n_traj, 4), device=dev, dtype=torch.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        sums.zero_()
        scorer.partial_sums(c2w, vt, n_traj, sums)
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
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    sampler.mark_end()
    launches = _lib.kernel_launches()
    ms = e0.elapsed_time(e1)
    # end to end through the public API (host poses in, host scores out)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        terms = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])

    # ---- roofline of the dominant kernel (field_forward_kernel), measured live on rank 0 ----
    roof = None
    if rank == 0:
        roof = field_kernel_roofline(torch, scorer, c2w, vt, n_traj)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_path_rays_per_s(min_seconds=10.0)
    rays_per_step = V * world * R * 2
    if rank == 0:
        out = {
            "metric": METRIC, "value": rays_per_step * args.steps / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 MLP operands / f32 accumulate+compositing / f64 scoring",
            "data": "synthetic",
            "config": {"workload": f"planner candidate-view batch (BASELINE.json configs[2]): {V} poses/GPU x "
                                   f"{W}x{H} rays x 2 ensemble members, 128^3 occ grid, 16-level hash NeRF, sem-num 29, "
                                   "max_samples 1024, render+score (pred-info)", "views_per_gpu": V,
                       "rays_per_step": rays_per_step, "ensemble": 2, "n_trajectories": n_traj,
                       "l2": "per-step working set (>2 GB) exceeds the 126 MB L2; no flush needed",
                       "mean_samples_per_ray": roof.pop("_samples_per_ray") if roof else None,
                       "parallelism": f"views sharded over {world} rank(s), one all-reduce of [n_traj,4] f64"},
            "clocks": clocks,
            "e2e": {"value": rays_per_step * args.steps / (e2e_ms * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": int(V * (12 * 4 + 4)), "d2h_bytes_per_step": int(n_traj * 4 * 8)},
            "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "scores_sample": np.round(terms[0], 6).tolist(),
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def field_kernel_roofline(torch, scorer, c2w, vt, n_traj):
    """One instrumented pass: CUDA events around every field_forward launch (on the launching
    stream) + the per-launch sample counts -> achieved algorithmic GB/s of the gather."""
    from apnerf import _lib

    peak, peak_kind = _peaks()
    evs, counts = [], []
    r = scorer.renderer

    def hook(pc):
        if not pc.name.startswith("apnerf_field_forward"):
            return pc.invoke()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        pc.invoke()
        a1.record()
        evs.append((a0, a1))
        cnt = pc.keep[0].new_empty(0)  # locate the renderer this call belongs to through its counters tensor
        for rr in scorer.all_renderers():
            if any(t.data_ptr() == rr.counters[2:3].data_ptr() for t in pc.keep):
                cnt = rr.counters[6:7].clone() if pc.name.endswith("fused") else rr.counters[2:3].clone()
        counts.append(cnt)

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
    k_ms = sum(a.elapsed_time(b) for a, b in evs)
    n_samples = int(torch.cat(counts).sum().item())
    n_launch = sum(1 for c in counts if int(c.item()) > 0)
    step_ms = t0.elapsed_time(t1)
    bytes_per_sample = 1024  # 16 levels x 8 corners x 4 features x 2 B
    achieved = n_samples * bytes_per_sample / (k_ms * 1e-3) / 1e9
    n_rays = c2w.shape[0] * scorer.rays_per_view * len(scorer.fields)
    r = None
    return {"bound": "hbm", "kernel": "field_forward_kernel (hash-grid gather + fused tcgen05 MLPs)",
            "achieved": achieved, "peak": peak, "peak_kind": peak_kind + " HBM copy GB/s (MEASURED_PEAKS.json)",
            "unit": "GB/s", "frac": achieved / peak,
            # DRAM bytes per launch from the ncu --set full capture of this kernel (profiles/r01_field_kernel.md:
            # 104 B of dram read+write per sample -- the 48 MB table is L2-resident) x samples per launch
            "traffic": NCU_DRAM_BYTES_PER_SAMPLE * n_samples / max(1, n_launch),
            "algorithmic_bytes_per_launch": bytes_per_sample * n_samples / max(1, n_launch),
            "algorithmic_bytes_per_sample": bytes_per_sample, "samples_per_step": n_samples,
            "launches_with_work": n_launch, "avg_launch_ms": k_ms / max(1, n_launch),
            "kernel_share_of_step": k_ms / step_ms, "gsamples_per_s": n_samples / (k_ms * 1e-3) / 1e9,
            "mlp_tflops": n_samples * 81920 / (k_ms * 1e-3) / 1e12, "_samples_per_ray": n_samples / n_rays}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--views-per-gpu", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--concurrent-batches", type=int, default=1,
                    help="view batches rendered concurrently per GPU (each x the ensemble members, own streams)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
