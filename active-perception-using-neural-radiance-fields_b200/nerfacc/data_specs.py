"""Return types of ``traverse_grids`` -- same fields as the reference dataclasses
(perception/nerfacc/nerfacc/data_specs.py:12-87, 90-180)."""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class RaySamples:
    vals: torch.Tensor
    packed_info: Optional[torch.Tensor] = None
    ray_indices: Optional[torch.Tensor] = None
    is_valid: Optional[torch.Tensor] = None

    @property
    def device(self) -> torch.device:
        return self.vals.device


@dataclass
class RayIntervals:
    vals: torch.Tensor
    packed_info: Optional[torch.Tensor] = None
    ray_indices: Optional[torch.Tensor] = None
    is_left: Optional[torch.Tensor] = None
    is_right: Optional[torch.Tensor] = None

    @property
    def device(self) -> torch.device:
        return self.vals.device
