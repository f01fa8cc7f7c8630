"""Volume rendering ops with the reference's signatures (perception/nerfacc/nerfacc/volrend.py).

Packed (flattened) inputs run the fused CUDA kernels behind the C-ABI; batched
``(n_rays, n_samples)`` inputs keep the reference's plain tensor formulas (they are not on
the pipeline's path).  The alpha-based variants are thin wrappers.
"""
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor

from .._lib import call, require_cuda
from .pack import pack_info
from .scan import exclusive_sum


class _WeightsFromDensity(torch.autograd.Function):
    """(weights, trans, alphas) = f(sigmas[, prefix_trans]) on packed samples, one fused kernel
    each way.  Not differentiable w.r.t. t_starts / t_ends (as documented by the reference,
    volrend.py:39-41)."""

    @staticmethod
    def forward(ctx, t_starts, t_ends, sigmas, packed_info, prefix_trans):
        require_cuda(t_starts, t_ends, sigmas, packed_info, prefix_trans)
        # the kernels read float32 / int64 device pointers: normalise here (an fp16 sigma from an autocast field, say)
        t_starts, t_ends, sigmas = (t.float().contiguous() for t in (t_starts, t_ends, sigmas))
        chunk_starts = packed_info[:, 0].to(torch.int64).contiguous()
        chunk_cnts = packed_info[:, 1].to(torch.int64).contiguous()
        if prefix_trans is not None:
            prefix_trans = prefix_trans.float().contiguous()
        weights = torch.empty_like(sigmas)
        trans = torch.empty_like(sigmas)
        alphas = torch.empty_like(sigmas)
        if sigmas.numel():
            with torch.cuda.device(sigmas.device):
                call("apnerf_weights_from_density", chunk_cnts.numel(), chunk_starts, chunk_cnts, sigmas.numel(),
                     t_starts, t_ends, sigmas, prefix_trans, weights, trans, alphas)
        ctx.has_prefix = prefix_trans is not None
        ctx.save_for_backward(t_starts, t_ends, sigmas, chunk_starts, chunk_cnts,
                              prefix_trans if prefix_trans is not None else sigmas)
        return weights, trans, alphas

    @staticmethod
    def backward(ctx, g_weights, g_trans, g_alphas):
        t_starts, t_ends, sigmas, chunk_starts, chunk_cnts, prefix = ctx.saved_tensors
        prefix = prefix if ctx.has_prefix else None
        g_sigmas = torch.zeros_like(sigmas)
        g_prefix = torch.zeros_like(sigmas) if (ctx.has_prefix and ctx.needs_input_grad[4]) else None
        if sigmas.numel():
            with torch.cuda.device(sigmas.device):
                call("apnerf_weights_from_density_bwd", chunk_cnts.numel(), chunk_starts, chunk_cnts,
                     sigmas.numel(), t_starts, t_ends, sigmas, prefix,
                     None if g_weights is None else g_weights.contiguous(),
                     None if g_trans is None else g_trans.contiguous(),
                     None if g_alphas is None else g_alphas.contiguous(), g_sigmas, g_prefix)
        return None, None, g_sigmas, None, g_prefix


class _Accumulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, values, ray_indices, n_rays):
        require_cuda(weights, values, ray_indices)
        weights = weights.float().contiguous()
        ray_indices = ray_indices.to(torch.int64).contiguous()
        D = 1 if values is None else values.shape[-1]
        if values is not None:
            values = values.float().contiguous()
        outputs = torch.zeros((n_rays, D), device=weights.device, dtype=weights.dtype)
        if weights.numel():
            with torch.cuda.device(weights.device):
                call("apnerf_accumulate_along_rays", weights.numel(), D, weights, values, ray_indices, outputs)
        ctx.D = D
        ctx.has_values = values is not None
        ctx.save_for_backward(weights, values if values is not None else weights, ray_indices)
        return outputs

    @staticmethod
    def backward(ctx, g_out):
        weights, values, ray_indices = ctx.saved_tensors
        values = values if ctx.has_values else None
        g_out = g_out.contiguous()
        g_w = torch.zeros_like(weights) if ctx.needs_input_grad[0] else None
        g_v = torch.zeros_like(values) if (ctx.has_values and ctx.needs_input_grad[1]) else None
        if weights.numel() and (g_w is not None or g_v is not None):
            with torch.cuda.device(weights.device):
                call("apnerf_accumulate_along_rays_bwd", weights.numel(), ctx.D, weights, values, ray_indices,
                     g_out, g_w, g_v)
        return g_w, g_v, None, None


def render_transmittance_from_density(
    t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None, prefix_trans: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor]:
    if ray_indices is not None and packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    if packed_info is None:  # batched (n_rays, n_samples)
        sigmas_dt = sigmas * (t_ends - t_starts)
        alphas = 1.0 - torch.exp(-sigmas_dt)
        trans = torch.exp(-exclusive_sum(sigmas_dt))
        if prefix_trans is not None:
            trans = trans * prefix_trans
        return trans, alphas
    _, trans, alphas = _WeightsFromDensity.apply(t_starts, t_ends, sigmas, packed_info, prefix_trans)
    return trans, alphas


def render_weight_from_density(
    t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None, prefix_trans: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    if ray_indices is not None and packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    if packed_info is None:
        trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, prefix_trans=prefix_trans)
        return trans * alphas, trans, alphas
    return _WeightsFromDensity.apply(t_starts, t_ends, sigmas, packed_info, prefix_trans)


@torch.no_grad()
def render_visibility_from_density(
    t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None, early_stop_eps: float = 1e-4,
    alpha_thre: float = 0.0, prefix_trans: Optional[Tensor] = None,
) -> Tensor:
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays,
                                                      prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


def render_transmittance_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """T_i = prod_{j<i}(1 - alpha_j) (volrend.py:164-209) via the packed sum in log space."""
    if ray_indices is not None and packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    trans = torch.exp(exclusive_sum(torch.log1p(-alphas), packed_info))
    if prefix_trans is not None:
        trans = trans * prefix_trans
    return trans


def render_weight_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    trans = render_transmittance_from_alpha(alphas, packed_info, ray_indices, n_rays, prefix_trans)
    return trans * alphas, trans


@torch.no_grad()
def render_visibility_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, early_stop_eps=1e-4,
                                 alpha_thre=0.0, prefix_trans=None):
    trans = render_transmittance_from_alpha(alphas, packed_info, ray_indices, n_rays, prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor] = None, ray_indices: Optional[Tensor] = None,
                          n_rays: Optional[int] = None) -> Tensor:
    if values is not None:
        assert values.dim() == weights.dim() + 1
        assert weights.shape == values.shape[:-1]
    if ray_indices is not None:
        assert n_rays is not None, "n_rays must be provided"
        assert weights.dim() == 1, "weights must be flattened"
        return _Accumulate.apply(weights, values, ray_indices, int(n_rays))
    src = weights[..., None] if values is None else weights[..., None] * values
    return torch.sum(src, dim=-2)


def accumulate_along_rays_(weights: Tensor, values: Optional[Tensor] = None, ray_indices: Optional[Tensor] = None,
                           outputs: Optional[Tensor] = None) -> None:
    """In-place variant (volrend.py:553-576); no autograd, like index_add_ on a leaf buffer."""
    if values is not None:
        assert values.dim() == weights.dim() + 1
        assert weights.shape == values.shape[:-1]
    if ray_indices is not None:
        assert weights.dim() == 1, "weights must be flattened"
        D = 1 if values is None else values.shape[-1]
        assert outputs.dim() == 2 and outputs.shape[-1] == D, "outputs must be of shape (n_rays, D)"
        assert outputs.is_contiguous()
        require_cuda(weights, values, ray_indices, outputs)
        if weights.numel():
            with torch.cuda.device(weights.device):
                assert outputs.dtype == torch.float32, "outputs must be float32"
                call("apnerf_accumulate_along_rays", weights.numel(), D, weights.detach().float().contiguous(),
                     None if values is None else values.detach().float().contiguous(),
                     ray_indices.to(torch.int64).contiguous(), outputs)
    else:
        src = weights[..., None] if values is None else weights[..., None] * values
        outputs.add_(src.sum(dim=-2))


def rendering(
    t_starts: Tensor, t_ends: Tensor, ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
    rgb_sigma_fn: Optional[Callable] = None, rgb_alpha_fn: Optional[Callable] = None,
    render_bkgd: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor, Dict]:
    """volrend.py:17-161."""
    if ray_indices is not None:
        assert t_starts.shape == t_ends.shape == ray_indices.shape, \
            "Since nerfacc 0.5.0, t_starts, t_ends and ray_indices must have the same shape (N,). "
    if rgb_sigma_fn is None and rgb_alpha_fn is None:
        raise ValueError("At least one of `rgb_sigma_fn` and `rgb_alpha_fn` should be specified.")
    if rgb_sigma_fn is not None:
        if t_starts.shape[0] != 0:
            rgbs, sigmas = rgb_sigma_fn(t_starts, t_ends, ray_indices)
        else:
            rgbs = torch.empty((0, 3), device=t_starts.device)
            sigmas = torch.empty((0,), device=t_starts.device)
        assert rgbs.shape[-1] == 3, "rgbs must have 3 channels, got {}".format(rgbs.shape)
        assert sigmas.shape == t_starts.shape, "sigmas must have shape of (N,)! Got {}".format(sigmas.shape)
        weights, trans, alphas = render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices,
                                                            n_rays=n_rays)
        extras = {"weights": weights, "alphas": alphas, "trans": trans, "sigmas": sigmas, "rgbs": rgbs}
    else:
        if t_starts.shape[0] != 0:
            rgbs, alphas = rgb_alpha_fn(t_starts, t_ends, ray_indices)
        else:
            rgbs = torch.empty((0, 3), device=t_starts.device)
            alphas = torch.empty((0,), device=t_starts.device)
        assert rgbs.shape[-1] == 3, "rgbs must have 3 channels, got {}".format(rgbs.shape)
        assert alphas.shape == t_starts.shape, "alphas must have shape of (N,)! Got {}".format(alphas.shape)
        weights, trans = render_weight_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays)
        extras = {"weights": weights, "trans": trans, "rgbs": rgbs, "alphas": alphas}
    colors = accumulate_along_rays(weights, values=rgbs, ray_indices=ray_indices, n_rays=n_rays)
    opacities = accumulate_along_rays(weights, values=None, ray_indices=ray_indices, n_rays=n_rays)
    depths = accumulate_along_rays(weights, values=(t_starts + t_ends)[..., None] / 2.0, ray_indices=ray_indices,
                                   n_rays=n_rays)
    depths = depths / opacities.clamp_min(torch.finfo(rgbs.dtype).eps)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opacities)
    return colors, opacities, depths, extras
