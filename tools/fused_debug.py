import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apnerf
from apnerf import synthetic, _lib
dev = "cuda:0"
est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
est.binaries = synthetic.make_occupancy(128, seed=1); est = est.to(dev).eval()
f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=29)
f = synthetic.init_trained_like(f, seed=2).to(dev).eval()
V, W, H = 16, 320, 240
c2w = torch.from_numpy(apnerf.scoring.poses_to_c2w(synthetic.make_poses_corridor(V, seed=3))).to(dev)
ro = torch.empty((V*W*H, 3), device=dev); rd = torch.empty_like(ro)
_lib.call("apnerf_generate_rays", V, c2w, W, H, 160.0, W*H, None, ro, rd)
opts = dict(near_plane=0.1, render_step_size=1e-3, cone_angle=0.004, alpha_thre=0.01)
r = apnerf.FusedRenderer(dev, 29)
stats = []
def hook(it, rr):
    c = rr.counters.cpu().numpy()
    stats.append((it, int(c[0]), int(c[2]), int(c[6])))
for fuse in (True, False):
    stats.clear()
    r.render(f, est, ro, rd, W*H, fuse_compositor=fuse, debug_hook=hook, poll_every=0, max_samples=64, **opts)
    torch.cuda.synchronize()
    print("fuse", fuse, [s for s in stats[:16]])
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r.render(f, est, ro, rd, W*H, fuse_compositor=fuse, **opts)
        torch.cuda.synchronize(); print("  render ms", 1e3*(time.perf_counter()-t0))
# time kernels individually for one mid iteration using events
orig = _lib.LIB.call
times = {}
def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *a); e1.record()
    times.setdefault(name, []).append((e0, e1))
import importlib
rmod = sys.modules[[m for m in sys.modules if m.endswith(".render")][0]]
for fuse in (True, False):
    times.clear(); rmod.call = timed
    r.render(f, est, ro, rd, W*H, fuse_compositor=fuse, **opts); torch.cuda.synchronize(); rmod.call = orig
    print("fuse", fuse, {k: round(sum(a.elapsed_time(b) for a, b in v), 2) for k, v in times.items()})
