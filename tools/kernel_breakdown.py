"""Live (warm, unprofiled) per-kernel breakdown of one bench step: CUDA events around EVERY C-ABI launch of a
sequential pass (members rendered one after the other), plus per-launch (samples, ms) of the field kernel.
Usage on the GPU box:  python tools/kernel_breakdown.py [views]"""
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apnerf  # noqa: E402
import bench  # noqa: E402
from apnerf import _lib, synthetic  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
est.binaries = synthetic.make_occupancy(128, seed=1)
est = est.to(dev).eval()
fields = [synthetic.init_trained_like(apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=bench.N_SEM),
                                      seed=s).to(dev).eval() for s in (2, 12)]
scorer = apnerf.PredictiveInformationScorer(fields, [est, est], bench.W, bench.H, bench.HFOV_FOCAL, device=dev,
                                            views_per_batch=V, **bench.OPTS)
c2w = torch.from_numpy(apnerf.scoring.poses_to_c2w(synthetic.make_poses(V, seed=3))).to(dev)
vt = torch.zeros(V, dtype=torch.int32, device=dev)
for _ in range(3):
    scorer.partial_sums(c2w, vt, 1)
torch.cuda.synchronize()

evs = []


def hook(pc):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    pc.invoke()
    a1.record()
    cnt = None
    if pc.name.startswith("apnerf_field_forward"):
        for rr in scorer.all_renderers():
            if hasattr(rr, "counters") and any(t.data_ptr() == rr.counters[2:3].data_ptr() for t in pc.keep):
                cnt = torch.stack([rr.counters[2], rr.counters[0]]).clone()  # rows this iteration, live rays
    evs.append((pc.name, a0, a1, cnt))


_lib.CALL_HOOK = hook
scorer.interleave = False
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
scorer.partial_sums(c2w, vt, 1)
t1.record()
torch.cuda.synchronize()
_lib.CALL_HOOK = None
tot = collections.defaultdict(float)
n = collections.Counter()
for name, a, b, _ in evs:
    tot[name] += a.elapsed_time(b)
    n[name] += 1
step = t0.elapsed_time(t1)
print(f"sequential instrumented step {step:.2f} ms, {len(evs)} hooked launches, sum of kernels {sum(tot.values()):.2f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {k:34s} {n[k]:4d} launches {v:8.3f} ms  {100 * v / step:5.1f} %")
print("field launches of member 0 (rows, live rays, ms, G rows/s):")
rows = [(int(c[0]), int(c[1]), a.elapsed_time(b)) for name, a, b, c in evs if c is not None]
for r, l, ms in rows[: len(rows) // 2]:
    print(f"  {r:9d} {l:8d} {ms:8.3f} {r / ms / 1e6 if ms > 0 else 0:6.2f}")
for key in ("apnerf_render_march", "apnerf_render_composite"):
    series = [a.elapsed_time(b) for name, a, b, _ in evs if name == key]
    half = series[: len(series) // 2]
    print(key, "ms per launch, member 0 (every 4th launch after the 24th):", " ".join(f"{v:.3f}" for v in half[:24]), "|",
          " ".join(f"{v:.3f}" for v in half[24::4]))
scorer.interleave = True
t0.record()
for _ in range(5):
    scorer.partial_sums(c2w, vt, 1)
t1.record()
torch.cuda.synchronize()
print(f"interleaved step {t0.elapsed_time(t1) / 5:.2f} ms")
