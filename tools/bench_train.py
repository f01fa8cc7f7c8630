"""Secondary measurement (BASELINE.json configs[3]): NeRF training step on synthetic posed rays,
8192 rays per batch per GPU, data-parallel gradient all-reduce over NCCL when launched with torchrun.
Prints one JSON line (rank 0).  Not the driver's bench (bench.py is); kept for DESIGN.md / profiles."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apnerf  # noqa: E402
from apnerf import synthetic, training  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N_RAYS, STEPS, WARM = 8192, int(os.environ.get("STEPS", 30)), 5
est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
est.binaries = synthetic.make_occupancy(128, seed=1)
est.occs = est.binaries.flatten().float() * 0.5
est = est.to(dev)
f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=29)
f = synthetic.init_trained_like(f, seed=2, density_gain=2.0).to(dev)
opt = torch.optim.Adam(f.parameters(), lr=1e-3, eps=1e-15)  # pipeline.py:173-178
g = torch.Generator().manual_seed(4 + rank)
d = torch.randn((N_RAYS, 3), generator=g)
d = d / d.norm(dim=-1, keepdim=True)
batch = dict(rays=apnerf.Rays(origins=torch.tensor([0.1, 1.5, -0.2]).expand(N_RAYS, 3).contiguous().to(dev), viewdirs=d.to(dev)),
             pixels=torch.rand((N_RAYS, 3), generator=g).to(dev), dep=(torch.rand(N_RAYS, generator=g) * 4 + 0.5).to(dev),
             sem=torch.randint(0, 29, (N_RAYS,), generator=g).to(dev), color_bkgd=torch.rand(3, generator=g).to(dev))
n_samples = 0
for i in range(WARM + STEPS):
    if i == WARM:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
    out = training.training_step(f, est, opt, batch, step=1000 + i, update_occupancy=(i % 16 == 0))
    if i >= WARM and out:
        n_samples += out["n_samples"]
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"metric": "training rays/s (config 4)", "value": N_RAYS * world * STEPS / dt, "unit": "rays/s",
                      "n_gpus": world, "ms_per_step": 1e3 * dt / STEPS, "rays_per_batch_per_gpu": N_RAYS,
                      "mean_samples_per_step": n_samples / STEPS, "loss": out["loss"] if out else None,
                      "note": "fused tcgen05 forward (activations saved) + tcgen05 backward chain + hash-grid scatter + "
                              "packed volrend CUDA fwd/bwd; weight gradients dW = dY^T X as library GEMMs"}))
if world > 1:
    dist.destroy_process_group()
