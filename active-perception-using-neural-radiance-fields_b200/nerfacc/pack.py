"""pack_info with the reference signature (perception/nerfacc/nerfacc/pack.py:10-49)."""
from typing import Optional

import torch
from torch import Tensor

from .._lib import LIB, call


@torch.no_grad()
def pack_info(ray_indices: Tensor, n_rays: Optional[int] = None) -> Tensor:
    assert ray_indices.dim() == 1, "ray_indices must be a 1D tensor with shape (n_samples)."
    if not ray_indices.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1
    ray_indices = ray_indices.contiguous().to(torch.int64)
    dev = ray_indices.device
    packed_info = torch.empty((n_rays, 2), device=dev, dtype=torch.int64)
    n_scr = 2 * n_rays + int(LIB.raw("apnerf_scan_scratch_elems")(n_rays))
    scratch = torch.empty(n_scr, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        call("apnerf_pack_info", ray_indices.numel(), ray_indices, int(n_rays), packed_info, scratch)
    return packed_info
