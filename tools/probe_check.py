"""Scheduler check on ONE GPU: (a) cost and accuracy of the probe render behind PredictiveInformationScorer.view_cost_proxy
against the true per-view field rows of the full-resolution render, (b) how well a split into N shards balances in TIME
when the shards are rendered one after the other on this GPU: contiguous slices vs LPT on the probe's cost.
Usage: python tools/probe_check.py [n_shards]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apnerf  # noqa: E402
import bench  # noqa: E402
from apnerf import synthetic  # noqa: E402
from apnerf.scoring import lpt_assign, shard_range  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
V = 256
dev = torch.device("cuda", 0)
est, fields = bench._scene(apnerf, synthetic, dev, 6.0)
scorer = apnerf.PredictiveInformationScorer(fields, [est, est], bench.W, bench.H, bench.HFOV_FOCAL, device=dev, **bench.OPTS)
c2w = torch.from_numpy(apnerf.scoring.poses_to_c2w(synthetic.make_poses(V, seed=3))).to(dev)
vt = torch.zeros(V, dtype=torch.int32, device=dev)

for _ in range(2):
    cost = scorer.view_cost_proxy(c2w)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    cost = scorer.view_cost_proxy(c2w)
torch.cuda.synchronize()
print(f"probe: {scorer._probe['k']} rays/view, min_samples {scorer.probe_min_samples}, {scorer.probe_iters} iterations: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call "
      f"for {V} views; cost min/median/max {cost.min():.0f} {np.median(cost):.0f} {cost.max():.0f}")

# true rows per view: full-resolution renders with call_rows, 32 views at a time, both members
R = bench.W * bench.H
true_rows = np.zeros(V)
r = apnerf.FusedRenderer(dev, bench.N_SEM)
for v0 in range(0, V, 32):
    nv = min(32, V - v0)
    rays_o, rays_d = torch.empty((nv * R, 3), device=dev), torch.empty((nv * R, 3), device=dev)
    apnerf._lib.call("apnerf_generate_rays", nv, c2w[v0:v0 + nv].contiguous(), bench.W, bench.H, bench.HFOV_FOCAL, R, None, rays_o, rays_d)
    for f in fields:
        cr = torch.zeros(nv, device=dev, dtype=torch.int32)
        r.render(f, est, rays_o, rays_d, R, max_samples=1024, call_rows=cr, **bench.OPTS)
        true_rows[v0:v0 + nv] += cr.cpu().numpy()
print(f"true rows per view: min/median/max {true_rows.min():.0f} {np.median(true_rows):.0f} {true_rows.max():.0f}, total {true_rows.sum():.3e}")
print(f"correlation(probe cost, true rows) = {np.corrcoef(cost, true_rows)[0, 1]:.4f};  "
      f"probe/true scale {cost.sum() / true_rows.sum():.5f}, relative error of the scaled probe: "
      f"median {np.median(np.abs(cost / cost.sum() * true_rows.sum() - true_rows) / true_rows):.3f}")


from apnerf.scoring import _PassQueue  # noqa: E402


def _queue(order, plan_cost):
    q = _PassQueue(dev)
    q.add_local(scorer.plan_batches(np.asarray(order), plan_cost))
    return q


def time_shards(shards, label, plan_cost=None, min_batches=3):
    """Render every shard alone on this GPU (3 timed repeats each).  plan_cost: per-view costs handed to the scorer's
    pass planner (heaviest-first order, equal-cost passes); None: one even pass per 64 views."""
    ms, rows = [], []
    scorer.min_batches = min_batches
    for idx in shards:
        idx = np.asarray(idx)
        sel = torch.from_numpy(idx).to(dev)
        cs, vs = c2w.index_select(0, sel).contiguous(), vt.index_select(0, sel).contiguous()
        if plan_cost is not None:
            local = np.asarray(plan_cost)[idx]
            scorer.schedule = lambda c, pg, local=local: _queue(np.argsort(-local, kind="stable"), local)
        else:
            scorer.schedule = lambda c, pg, n=len(idx): _queue(np.arange(n), None)
        for _ in range(2):
            scorer.partial_sums(cs, vs, 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            scorer.partial_sums(cs, vs, 1)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b) / 3)
        rows.append(true_rows[idx].sum())
    ms, rows = np.asarray(ms), np.asarray(rows)
    print(f"{label}: shard ms {np.round(ms, 1).tolist()}  max {ms.max():.1f} mean {ms.mean():.1f} max/mean {ms.max() / ms.mean():.3f};  "
          f"rows max/mean {rows.max() / rows.mean():.3f}")
    return ms


contig = [np.arange(*shard_range(V, k, N)) for k in range(N)]
lpt_probe, lpt_true = lpt_assign(cost, N), lpt_assign(true_rows, N)
if os.environ.get("PROBE_CHECK_QUICK"):
    time_shards(lpt_probe, f"LPT(probe) x{N}, one pass")
    time_shards(lpt_probe, f"LPT(probe) x{N}, equal-cost passes (min 3)", plan_cost=cost, min_batches=3)
    sys.exit(0)
time_shards(contig, f"contiguous x{N}, one pass")
time_shards(lpt_probe, f"LPT(probe) x{N}, one pass")
for mb in (3, 4, 6):
    time_shards(lpt_probe, f"LPT(probe) x{N}, equal-cost passes (min {mb})", plan_cost=cost, min_batches=mb)
time_shards(lpt_true, f"LPT(true rows) x{N}, equal-cost passes (min 4)", plan_cost=true_rows, min_batches=4)
time_shards(contig, f"contiguous x{N}, equal-cost passes (min 4)", plan_cost=cost, min_batches=4)
