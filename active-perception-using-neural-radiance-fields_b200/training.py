"""NeRF training step on posed rgb / depth / semantic rays -- the caller side of the path in training
(scripts/pipeline.py:447-532) with data-parallel gradient all-reduce.

Per step and model: occupancy-grid update every 16 steps (``estimator.update_every_n_steps``), the
train-mode render (``render_image_with_occgrid_with_depth_guide``), the reference's loss
``10 * smoothL1(rgb) + smoothL1(depth) / 5 + CE(sem) / 2`` (pipeline.py:507-511), NaN-gradient guard
(:520-529), optimizer step.  Multi-GPU: every rank draws its own rays; gradients of the three flat
parameter vectors (25.24 M fp32 values, 101 MB) are all-reduced (NCCL over NVLink) before the step.
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .nerfacc import DensityOccEvalFn
from .render import Rays, render_image_with_occgrid_with_depth_guide


def allreduce_gradients(module: torch.nn.Module, process_group=None) -> None:
    """Average parameter gradients over the ranks: one flattened all-reduce per parameter tensor
    (three large tensors here, so bucketing is already done by the tcnn-style flat layout)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(process_group)
    if world == 1:
        return
    for p in module.parameters():
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        if p.grad.numel():
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=process_group)
            p.grad.div_(world)


def nerf_loss(rgb, depth, sem, batch: Dict[str, torch.Tensor]):
    """pipeline.py:507-511."""
    loss = F.smooth_l1_loss(rgb, batch["pixels"]) * 10
    loss = loss + F.smooth_l1_loss(depth.reshape(-1), batch["dep"].reshape(-1)) / 5
    if sem is not None:
        loss = loss + F.cross_entropy(sem, batch["sem"]) / 2
    return loss


def training_step(radiance_field, estimator, optimizer, batch: Dict[str, torch.Tensor], step: int, *,
                  near_plane: float = 0.1, render_step_size: float = 1e-3, cone_angle: float = 0.004,
                  alpha_thre: float = 0.01, occ_thre: float = 1e-2, scheduler=None, process_group=None,
                  update_occupancy: bool = True) -> Optional[Dict[str, float]]:
    """One optimisation step; ``batch`` holds ``rays`` (Rays of [n,3]), ``pixels [n,3]``, ``dep [n]``,
    ``sem [n] int64`` and ``color_bkgd [3]`` (habitat_to_data.py:205-272).  Returns the logged scalars,
    or None when the step was skipped (no samples, or NaN gradients)."""
    radiance_field.train()
    estimator.train()
    if update_occupancy:
        # pipeline.py:376-378 as a recognisable object: the estimator fuses the whole per-level update into
        # one launch of the field kernel (apnerf_occ_update)
        estimator.update_every_n_steps(step=step, occ_eval_fn=DensityOccEvalFn(radiance_field, render_step_size),
                                       occ_thre=occ_thre)
    out = render_image_with_occgrid_with_depth_guide(
        radiance_field, estimator, batch["rays"], near_plane=near_plane, render_step_size=render_step_size,
        render_bkgd=batch.get("color_bkgd"), cone_angle=cone_angle, alpha_thre=alpha_thre, depth=batch.get("dep"))
    if radiance_field.num_semantic_classes > 0:
        rgb, acc, depth, sem, n_samples = out
    else:
        (rgb, acc, depth, n_samples), sem = out, None
    if n_samples == 0:  # pipeline.py:491
        return None
    loss = nerf_loss(rgb, depth, sem, batch)
    optimizer.zero_grad()
    loss.backward()
    allreduce_gradients(radiance_field, process_group)
    for p in radiance_field.parameters():  # pipeline.py:520-529
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        if torch.isnan(p.grad).any():
            optimizer.zero_grad()
            return None
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return {"loss": float(loss.detach()), "n_samples": int(n_samples)}
