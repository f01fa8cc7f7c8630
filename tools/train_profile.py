"""Kernel-level breakdown of the training step (torch.profiler, CUDA activities).  GPU box only."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apnerf  # noqa: E402
from apnerf import synthetic, training  # noqa: E402

dev = torch.device("cuda", 0)
N = 8192
est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
est.binaries = synthetic.make_occupancy(128, seed=1)
est.occs = est.binaries.flatten().float() * 0.5
est = est.to(dev)
f = synthetic.init_trained_like(apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=29), seed=2,
                                density_gain=2.0).to(dev)
opt = torch.optim.Adam(f.parameters(), lr=1e-3, eps=1e-15)
g = torch.Generator().manual_seed(4)
d = torch.randn((N, 3), generator=g)
d = d / d.norm(dim=-1, keepdim=True)
batch = dict(rays=apnerf.Rays(origins=torch.tensor([0.1, 1.5, -0.2]).expand(N, 3).contiguous().to(dev), viewdirs=d.to(dev)),
             pixels=torch.rand((N, 3), generator=g).to(dev), dep=(torch.rand(N, generator=g) * 4 + 0.5).to(dev),
             sem=torch.randint(0, 29, (N,), generator=g).to(dev), color_bkgd=torch.rand(3, generator=g).to(dev))
for i in range(5):
    training.training_step(f, est, opt, batch, step=1001 + i, update_occupancy=False)
torch.cuda.synchronize()
STEPS = 10
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(STEPS):
        training.training_step(f, est, opt, batch, step=1001 + i, update_occupancy=False)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / STEPS, e.count / STEPS) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows if not r[0].startswith(("aten::", "autograd::", "Optimizer", "_Fused", "_Weights", "_Accum", "_Packed", "_TruncExp")))
print("device time per step (us), kernels only: %.0f" % tot)
for k, t, c in rows[:45]:
    print("%9.1f us  x%5.1f  %s" % (t, c, k[:110]))

print("\nhost side (self CPU time per step, us):")
rows = [(e.key, e.self_cpu_time_total / STEPS, e.count / STEPS) for e in prof.key_averages()]
rows.sort(key=lambda r: -r[1])
print("total self CPU per step: %.0f us" % sum(r[1] for r in rows))
for k, t, c in rows[:28]:
    print("%9.1f us  x%5.1f  %s" % (t, c, k[:100]))
