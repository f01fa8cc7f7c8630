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


def allreduce_gradients(module: torch.nn.Module, process_group=None, contributed: bool = True) -> torch.Tensor:
    """Sum parameter gradients over the ranks and divide by the number of ranks that CONTRIBUTED a batch (one
    flattened all-reduce per parameter tensor: three large tensors here, so bucketing is already done by the
    tcnn-style flat layout).  Every rank must call this every step, including a rank whose batch produced no
    samples (it contributes zeros and ``contributed=False``): the skip decision is taken from the reduced count, so
    the ranks stay in lock step.  Returns the number of contributing ranks as a 1-element DEVICE tensor -- nothing is
    read back here, so the host keeps running ahead of the GPU (a ``.item()`` at this point cost 3 ms per step on 8
    GPUs); the caller folds ``count == 0`` into the one host read it does anyway (the NaN guard)."""
    import torch.distributed as dist

    for p in module.parameters():
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    dev = next(module.parameters()).device
    flag = torch.tensor([1.0 if contributed else 0.0], device=dev)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return flag
    dist.all_reduce(flag, op=dist.ReduceOp.SUM, group=process_group)
    inv = 1.0 / flag.clamp_min(1.0)
    for p in module.parameters():
        if p.grad.numel():
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=process_group)
            p.grad.mul_(inv)
    return flag


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
                  update_occupancy: bool = True, read_loss: bool = True) -> Optional[Dict[str, float]]:
    """One optimisation step; ``batch`` holds ``rays`` (Rays of [n,3]), ``pixels [n,3]``, ``dep [n]``,
    ``sem [n] int64`` and ``color_bkgd [3]`` (habitat_to_data.py:205-272).  Returns the logged scalars,
    or None when the step was skipped (no samples on any rank, or NaN gradients).  ``read_loss=False`` leaves the
    loss on the device (``"loss"`` is then a 0-d tensor): no host synchronisation for logging."""
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
    optimizer.zero_grad()
    loss = None
    if n_samples > 0:  # pipeline.py:491 skips an empty batch; under data parallelism the rank still joins the all-reduce
        loss = nerf_loss(rgb, depth, sem, batch)
        loss.backward()
    n_contrib = allreduce_gradients(radiance_field, process_group, contributed=n_samples > 0)
    # pipeline.py:520-529: skip the step when any gradient is NaN -- and when no rank had samples (:491).  One fused
    # reduction and ONE host read per step instead of a sync per parameter; the decision is the same on every rank
    # because it is taken after the all-reduce.
    bad = torch.stack([torch.isnan(p.grad).any() for p in radiance_field.parameters() if p.grad.numel()]).any()
    if bool(bad | (n_contrib[0] == 0)):
        optimizer.zero_grad()
        return None
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    if loss is None:
        return {"loss": float("nan"), "n_samples": 0}
    return {"loss": float(loss.detach()) if read_loss else loss.detach(), "n_samples": int(n_samples)}


class EnsembleTrainer:
    """``nerf_training``'s inner loop over the ensemble (scripts/pipeline.py:403-532): every call of ``step``
    trains each member-model once on its own freshly drawn ray batch."""

    def __init__(self, fields, estimators, optimizers, *, near_plane=0.1, render_step_size=1e-3, cone_angle=0.004,
                 alpha_thre=0.01, occ_thre=1e-2, schedulers=None, process_group=None):
        self.fields, self.estimators, self.optimizers = list(fields), list(estimators), list(optimizers)
        self.schedulers = list(schedulers) if schedulers is not None else [None] * len(self.fields)
        self.opts = dict(near_plane=near_plane, render_step_size=render_step_size, cone_angle=cone_angle,
                         alpha_thre=alpha_thre, occ_thre=occ_thre, process_group=process_group)
        self.samples_seen = 0
        self.steps_done = 0

    def step(self, fetch, step: int, read_loss: bool = False):
        """fetch() -> a training batch (called once per member, as the reference draws a new batch per model)."""
        logs = []
        for f, e, o, sch in zip(self.fields, self.estimators, self.optimizers, self.schedulers):
            out = training_step(f, e, o, fetch(), step, scheduler=sch, read_loss=read_loss, **self.opts)
            if out is not None:
                self.samples_seen += out["n_samples"]
                self.steps_done += 1
                logs.append(out["loss"] if read_loss else None)
            else:
                logs.append(None)
        return logs
