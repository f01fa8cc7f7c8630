"""Target of the ncu capture of the field kernel (tools/gpu_prof_final.sh): renders V views of the bench scene once
through both ensemble members, one after the other (no cross-stream interleave, so the k-th field launch is
well defined), and prints the sample rows of every field launch in launch order -> gpurun_out/field_rows.json.
Usage: python tools/field_profile_target.py [views]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apnerf  # noqa: E402
import bench  # noqa: E402
from apnerf import synthetic  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
est, fields = bench._scene(apnerf, synthetic, dev, 6.0)
poses = synthetic.make_poses(256, seed=3)[:V]
c2w = torch.from_numpy(apnerf.scoring.poses_to_c2w(poses)).to(dev)
R = bench.W * bench.H
rays_o, rays_d = torch.empty((V * R, 3), device=dev), torch.empty((V * R, 3), device=dev)
apnerf._lib.call("apnerf_generate_rays", V, c2w, bench.W, bench.H, bench.HFOV_FOCAL, R, None, rays_o, rays_d)
rows = []
for m, f in enumerate(fields):
    r = apnerf.FusedRenderer(dev, bench.N_SEM)
    per_iter = []
    r.render(f, est, rays_o, rays_d, R, max_samples=1024, poll_every=0,
             debug_hook=lambda it, rr: per_iter.append(rr.counters[2:3].clone()), **bench.OPTS)
    torch.cuda.synchronize()
    rows += [dict(member=m, iteration=i, rows=int(c.item())) for i, c in enumerate(per_iter)]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/field_rows.json", "w"))
print("field launches:", len(rows), "first rows:", [r["rows"] for r in rows[:8]])
