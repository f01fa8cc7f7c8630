"""Packed inclusive / exclusive sums with autograd, reference signatures
(perception/nerfacc/nerfacc/scan.py:15-97, 180-229).  Products (inclusive_prod /
exclusive_prod) are only reached from the alpha-based renderers, which the pipeline does not
use; they are provided through the exact identity with a scan in
log space only for batched inputs."""
from typing import Optional

import torch
from torch import Tensor

from .._lib import call, require_cuda


def _packed_sum(chunk_starts, chunk_cnts, inputs, inclusive, normalize, backward):
    require_cuda(chunk_starts, chunk_cnts, inputs)
    outputs = torch.empty_like(inputs)
    if inputs.numel() == 0:
        return outputs
    with torch.cuda.device(inputs.device):
        call("apnerf_packed_sum", chunk_cnts.numel(), chunk_starts, chunk_cnts, inputs.numel(), inputs, outputs,
             1 if inclusive else 0, 1 if normalize else 0, 1 if backward else 0)
    return outputs


class _PackedSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, chunk_starts, chunk_cnts, inputs, inclusive: bool, normalize: bool = False):
        chunk_starts = chunk_starts.contiguous()
        chunk_cnts = chunk_cnts.contiguous()
        inputs = inputs.contiguous()
        outputs = _packed_sum(chunk_starts, chunk_cnts, inputs, inclusive, normalize, False)
        if ctx.needs_input_grad[2]:
            ctx.inclusive, ctx.normalize = inclusive, normalize
            ctx.save_for_backward(chunk_starts, chunk_cnts)
        return outputs

    @staticmethod
    def backward(ctx, grad_outputs):
        grad_outputs = grad_outputs.contiguous()
        chunk_starts, chunk_cnts = ctx.saved_tensors
        assert ctx.normalize is False, "Only support backward for normalize==False."
        grad_inputs = _packed_sum(chunk_starts, chunk_cnts, grad_outputs, ctx.inclusive, False, True)
        return None, None, grad_inputs, None, None


def inclusive_sum(inputs: Tensor, packed_info: Optional[Tensor] = None) -> Tensor:
    if packed_info is None:
        return torch.cumsum(inputs, dim=-1)
    assert inputs.dim() == 1, "inputs must be flattened."
    assert packed_info.dim() == 2 and packed_info.shape[-1] == 2, "packed_info must be 2-D with shape (B, 2)."
    chunk_starts, chunk_cnts = packed_info.unbind(dim=-1)
    return _PackedSum.apply(chunk_starts, chunk_cnts, inputs, True, False)


def exclusive_sum(inputs: Tensor, packed_info: Optional[Tensor] = None) -> Tensor:
    if packed_info is None:
        return torch.cumsum(torch.cat([torch.zeros_like(inputs[..., :1]), inputs[..., :-1]], dim=-1), dim=-1)
    assert inputs.dim() == 1, "inputs must be flattened."
    assert packed_info.dim() == 2 and packed_info.shape[-1] == 2, "packed_info must be 2-D with shape (B, 2)."
    chunk_starts, chunk_cnts = packed_info.unbind(dim=-1)
    return _PackedSum.apply(chunk_starts, chunk_cnts, inputs, False, False)
