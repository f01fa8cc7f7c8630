"""NGPRadianceField without tiny-cuda-nn: same constructor, methods, attributes and
``state_dict`` keys as the reference (perception/models/radiance_fields/ngp.py:69-238), with the
hash grid + three MLPs evaluated by ONE fused sm_100a kernel (csrc/field_kernel.cuh).

Parameter layout (tcnn-compatible, flat fp32, so reference checkpoints load):
  mlp_base.params = [W1 (neurons x 64) | W2 (neurons x neurons) ... | W_out (16 x neurons) | grid (entries x 4)]
  mlp_head.params = [64 x 32 | 64 x 64 | 16 x 64]      (inputs padded to 32 with 1.0, outputs 3 -> 16)
  mlp_sem.params  = [64 x 16 | 64 x 64 | 32 x 64]      (inputs padded to 16 with 1.0, outputs C -> 32)
  direction_encoding.params = []  (SH has no parameters; kept so every named parameter exists)
All matrices are row-major [out, in], no biases.
"""
from typing import Optional, Callable, List, Union

import numpy as np
import torch

from .._lib import LIB, call, require_cuda


class _TruncExp(torch.autograd.Function):
    """exp forward, gradient clamped at exp(15) (ngp.py:23-39)."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, max=15))


trunc_exp = _TruncExp.apply


def hashgrid_levels(n_levels, base_resolution, max_resolution, log2_hashmap_size):
    """[n_levels, 5] uint32 rows {scale (f32 bits), resolution, entries, first entry, hashed} and
    the total number of table entries (tcnn GridEncoding; ngp.py:103-105,123-133)."""
    per_level_scale = np.exp((np.log(max_resolution) - np.log(base_resolution)) / (n_levels - 1))
    meta = np.zeros((n_levels, 5), dtype=np.uint32)
    offset = 0
    for lvl in range(n_levels):
        scale = np.exp2(lvl * np.log2(per_level_scale)) * base_resolution - 1.0
        res = int(np.ceil(scale)) + 1
        dense = res ** 3
        size = min((dense + 7) // 8 * 8, 1 << log2_hashmap_size)
        meta[lvl] = (np.float32(scale).view(np.uint32), res, size, offset, 1 if size < dense else 0)
        offset += size
    return meta, offset


class _FlatParams(torch.nn.Module):
    """Stand-in for a tcnn module: one flat fp32 ``params`` Parameter."""

    def __init__(self, n_params: int, n_output_dims: int = 0):
        super().__init__()
        self.params = torch.nn.Parameter(torch.zeros(n_params, dtype=torch.float32))
        self.n_output_dims = n_output_dims


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


def _xavier_(flat: torch.Tensor, dims, generator=None):
    o = 0
    for n_out, n_in in dims:
        bound = float(np.sqrt(6.0 / (n_in + n_out)))
        flat[o:o + n_out * n_in].uniform_(-bound, bound, generator=generator)
        o += n_out * n_in
    return o


def _umma_pack(w: torch.Tensor) -> torch.Tensor:
    """[N, K] fp16 -> UMMA K-major no-swizzle image: element (n, k) at (k//8)*(N*8) + n*8 + k%8."""
    n, k = w.shape
    return w.reshape(n, k // 8, 8).permute(1, 0, 2).contiguous().reshape(-1)


class _HashGridEncode(torch.autograd.Function):
    """fp16 multiresolution hash-grid features of aabb-normalised points, differentiable w.r.t. the
    fp32 table (CUDA forward ``apnerf_hashgrid_encode``, backward ``apnerf_hashgrid_encode_bwd``)."""

    @staticmethod
    def forward(ctx, x01, table_f32, meta, n_levels):
        import ctypes

        x01 = x01.contiguous()
        n = x01.shape[0]
        table_h = table_f32.detach().to(torch.float16).contiguous()
        enc = torch.zeros((n, 64), device=x01.device, dtype=torch.float16)
        out = enc if n_levels == 16 else torch.empty((n, n_levels * 4), device=x01.device, dtype=torch.float16)
        if n:
            with torch.cuda.device(x01.device):
                call("apnerf_hashgrid_encode", n, x01, n_levels, meta.ctypes.data_as(ctypes.c_void_p), table_h, out,
                     None)
        if out is not enc:
            enc[:, : n_levels * 4] = out
        ctx.meta, ctx.n_levels, ctx.table_shape = meta, n_levels, table_f32.shape
        ctx.save_for_backward(x01)
        return enc.float()  # values are fp16-representable; the graph stays fp32

    @staticmethod
    def backward(ctx, g):
        import ctypes

        (x01,) = ctx.saved_tensors
        n = x01.shape[0]
        d_table = torch.zeros(ctx.table_shape, device=g.device, dtype=torch.float32)
        g = g[:, : ctx.n_levels * 4].contiguous().float()
        if n:
            with torch.cuda.device(g.device):
                call("apnerf_hashgrid_encode_bwd", n, x01, ctx.n_levels, ctx.meta.ctypes.data_as(ctypes.c_void_p), g,
                     d_table)
        return None, d_table, None, None


LOSS_SCALE = 128.0  # tcnn's default loss scale for fp16 networks
# The weight-gradient GEMMs reduce over the samples in chunks of this many rows (buffers are padded with zero
# rows): the library then only ever sees [chunks, out, CHUNK] x [chunks, CHUNK, in] problems.  With the raw
# sample count as the GEMM's K every training step is a new problem shape, and planning one costs ~30 ms.
WGRAD_CHUNK = 32768


WGRAD_LIBRARY_GEMM = False  # True: the round-1 path (one torch.bmm), kept as the on-GPU cross-check of the kernel


def _wgrad_library(field, G, X, base_grad, head_grad, sem_grad):
    """dW through ONE library GEMM G^T . X (fp16 operands, fp32 accumulation), reduced over fixed-size chunks."""
    chunks = X.shape[0] // WGRAD_CHUNK
    GX = torch.bmm(G.view(chunks, WGRAD_CHUNK, _G_WIDTH).transpose(1, 2), X.view(chunks, WGRAD_CHUNK, _X_WIDTH),
                   out_dtype=torch.float32).sum(0) / LOSS_SCALE
    dW = [GX[_G_COLS[g][0]:_G_COLS[g][1], _X_COLS[x][0]:_X_COLS[x][1]].reshape(-1) for g, x in _WGRAD_BLOCKS]
    base_grad[: field._n_base_w] += torch.cat(dW[0:3])
    head_grad += torch.cat(dW[3:6])
    if sem_grad is not None:
        f = field._sem_dims_flat[2]
        sem_grad += torch.cat(dW[6:8] + [dW[8][: f[0] * f[1]]])


def _padded_rows(n: int) -> int:
    return max(1, (n + WGRAD_CHUNK - 1) // WGRAD_CHUNK) * WGRAD_CHUNK


def _rows(n: int, width: int, device, dtype=torch.float16) -> torch.Tensor:
    """[padded n, width] buffer whose rows >= n are zero (the kernels write rows < n)."""
    t = torch.empty((_padded_rows(n), width), device=device, dtype=dtype)
    t[n:].zero_()
    return t


# Column layout of the two wide fp16 matrices of the training path.  X holds the forward activations the
# backward needs (written by apnerf_field_forward_train), G the gradients w.r.t. every layer's output x LOSS_SCALE
# (written by apnerf_field_backward).  Keeping them as column slices of one matrix each makes ALL weight
# gradients one library GEMM  G^T . X  [576 x 624]; the nine diagonal blocks listed in _WGRAD_BLOCKS are dW.
_X_COLS = dict(enc=(0, 64), h1=(64, 192), h2=(192, 320), xh=(320, 352), xs=(352, 368), hh1=(368, 432), hh2=(432, 496),
               hs1=(496, 560), hs2=(560, 624))
_G_COLS = dict(g_h1=(0, 128), g_h2=(128, 256), g_base=(256, 272), g_hh1=(272, 336), g_hh2=(336, 400),
               g_out_h=(400, 416), g_hs1=(416, 480), g_hs2=(480, 544), g_out_s=(544, 576))
_X_WIDTH, _G_WIDTH = 624, 576
# (gradient block, activation block) in the order of the flat parameter vectors [W1|W2|W3], [WH1|WH2|WH3], [WS1|WS2|WS3]
_WGRAD_BLOCKS = (("g_h1", "enc"), ("g_h2", "h1"), ("g_base", "h2"), ("g_hh1", "xh"), ("g_hh2", "hh1"), ("g_out_h", "hh2"),
                 ("g_hs1", "xs"), ("g_hs2", "hs1"), ("g_out_s", "hs2"))


def _cols(mat: torch.Tensor, layout: dict, name: str):
    """Device pointer of a column block of a wide row-major fp16 matrix (the kernels take the row stride)."""
    import ctypes

    return ctypes.c_void_p(mat.data_ptr() + 2 * layout[name][0])


class _FusedMLPs(torch.autograd.Function):
    """The hash grid + three MLPs of the differentiable (training) path as two kernels:

    forward   ``apnerf_field_forward_train`` -- the fused inference kernel, instantiated to emit the raw fp16
              network outputs and to save every layer's activation;
    backward  ``apnerf_field_backward`` (tcgen05 chain of dX = dY . W with the ReLU masks, gradients x LOSS_SCALE
              in fp16 like tcnn) + ``apnerf_hashgrid_encode_bwd`` (scatter-add into the fp32 table gradient);
              the weight gradients dW = dY^T . X are one plain library GEMM over all samples (fp16 operands,
              fp32 accumulation and output), as tcnn computes them with CUTLASS.

    Inputs: positions, directions [n, 3] f32 and the flat fp32 parameter vectors (the SH one is empty).  Outputs: density
    logit [n], rgb logits [n, 3], semantic logits [n, C] (fp16 values upcast)."""

    @staticmethod
    def forward(ctx, pos, dirs, base_params, head_params, sem_params, dir_params, field):
        import ctypes

        n = pos.shape[0]
        dev = pos.device
        C = field.num_semantic_classes
        weights, table = field._packed()
        X = _rows(n, _X_WIDTH, dev)
        dens = torch.empty(n, device=dev, dtype=torch.float32)
        rgb = torch.empty((n, 3), device=dev, dtype=torch.float32)
        sem = torch.empty((n, C), device=dev, dtype=torch.float32) if C > 0 else None
        aabb_host = field.aabb_host()
        if n:
            with torch.cuda.device(dev):
                call("apnerf_field_forward_train", n, pos, dirs, aabb_host.ctypes.data_as(ctypes.c_void_p),
                     field.n_levels, field._meta.ctypes.data_as(ctypes.c_void_p), table, weights, dens, rgb, sem, C,
                     _X_WIDTH, *[_cols(X, _X_COLS, k) for k in ("enc", "h1", "h2", "xh", "xs", "hh1", "hh2", "hs1", "hs2")])
        ctx.field = field
        ctx.save_for_backward(pos, X)
        if C > 0:
            return dens, rgb, sem
        return dens, rgb

    @staticmethod
    def backward(ctx, g_dens, g_rgb, g_sem=None):
        import ctypes

        field = ctx.field
        pos, X = ctx.saved_tensors
        n = pos.shape[0]
        dev = pos.device
        C = field.num_semantic_classes
        g_dens = (torch.zeros(n, device=dev) if g_dens is None else g_dens.float()).contiguous()
        g_rgb = (torch.zeros((n, 3), device=dev) if g_rgb is None else g_rgb.float()).contiguous()
        if C > 0:
            g_sem = (torch.zeros((n, C), device=dev) if g_sem is None else g_sem.float()).contiguous()
        else:
            g_sem = None
        G = _rows(n, _G_WIDTH, dev)
        d_enc = torch.empty((n, 64), device=dev, dtype=torch.float32)
        # the flat gradient of mlp_base.params = [W1 | W2 | W3 | table]: the scatter kernel adds straight into its tail
        base_grad = torch.zeros(field._n_base_w + field._n_entries * 4, device=dev, dtype=torch.float32)
        d_table = base_grad[field._n_base_w:]
        if n:
            aabb_min, aabb_max = torch.split(field.aabb, 3, dim=-1)
            x01 = ((pos - aabb_min) / (aabb_max - aabb_min)).contiguous()
            with torch.cuda.device(dev):
                call("apnerf_field_backward", n, g_dens, g_rgb, g_sem, C, _X_WIDTH,
                     *[_cols(X, _X_COLS, k) for k in ("h1", "h2", "hh1", "hh2", "hs1", "hs2")],
                     field._packed_t(), float(LOSS_SCALE), _G_WIDTH,
                     *[_cols(G, _G_COLS, k) for k in ("g_out_h", "g_out_s", "g_hh2", "g_hs2", "g_hh1", "g_hs1", "g_base",
                                                      "g_h2", "g_h1")], d_enc)
                call("apnerf_hashgrid_encode_bwd", n, x01, field.n_levels,
                     field._meta.ctypes.data_as(ctypes.c_void_p),
                     d_enc[:, : field.n_levels * 4].contiguous(), d_table)
        # all nine weight gradients dW = G^T . X / LOSS_SCALE: tcgen05 split-K kernel, accumulators in TMEM
        # (csrc/field_wgrad_kernel.cuh); adds straight into the flat fp32 gradient vectors
        head_grad = torch.zeros(sum(a * b for a, b in field._head_dims), device=dev, dtype=torch.float32)
        sem_grad = torch.zeros(sum(a * b for a, b in field._sem_dims_flat), device=dev, dtype=torch.float32) if C > 0 else None
        if n:
            with torch.cuda.device(dev):
                if WGRAD_LIBRARY_GEMM:
                    _wgrad_library(field, G, X, base_grad, head_grad, sem_grad)
                else:
                    call("apnerf_field_wgrad", n, G, X, float(LOSS_SCALE), base_grad, head_grad, sem_grad,
                         field._sem_dims_flat[2][0] if C > 0 else 32)
        # the zero-length SH parameter vector gets a (zero-length) gradient too: the reference's NaN guard calls
        # torch.isnan(param.grad) on EVERY named parameter (scripts/pipeline.py:521-524)
        dir_grad = torch.zeros(0, device=dev, dtype=torch.float32)
        return None, None, base_grad, head_grad, sem_grad, dir_grad, None


class NGPRadianceField(torch.nn.Module):
    """Instant-NGP radiance field with an optional semantic head."""

    def __init__(
        self,
        aabb: Union[torch.Tensor, List[float]],
        num_dim: int = 3,
        use_viewdirs: bool = True,
        neurons: int = 128,
        layers: int = 4,
        density_activation: Optional[Callable] = None,
        unbounded: bool = False,
        base_resolution: int = 16,
        max_resolution: int = 4096,
        geo_feat_dim: int = 15,
        n_levels: int = 16,
        log2_hashmap_size: int = 19,
        num_semantic_classes: int = 0,
    ) -> None:
        super().__init__()
        if not isinstance(aabb, torch.Tensor):
            aabb = torch.tensor(aabb, dtype=torch.float32)
        self.register_buffer("aabb", aabb)
        self.num_dim = num_dim
        self.use_viewdirs = use_viewdirs
        # the sm_100a kernels (renderer, occupancy update, query_density) hard-code the reference's default
        # density activation trunc_exp(x - 1) (ngp.py:79): another callable would train with one density and render
        # with another, so it is refused instead of silently ignored
        if density_activation is not None:
            raise NotImplementedError("density_activation other than the default trunc_exp(x - 1) is not supported "
                                      "by the fused sm_100a kernels")
        self.density_activation = lambda x: trunc_exp(x - 1)
        self.unbounded = unbounded
        self.base_resolution = base_resolution
        self.max_resolution = max_resolution
        self.geo_feat_dim = geo_feat_dim
        self.n_levels = n_levels
        self.log2_hashmap_size = log2_hashmap_size
        self.num_semantic_classes = num_semantic_classes
        self.neurons = neurons
        self.layers = layers
        # what the sm_100a kernel is built for (the pipeline's configuration, config_*.yaml:17-18)
        if (num_dim != 3 or not use_viewdirs or unbounded or neurons != 128 or layers != 2 or geo_feat_dim != 15
                or n_levels > 16 or num_semantic_classes > 32):
            raise NotImplementedError(
                "the fused sm_100a field kernel supports num_dim=3, use_viewdirs=True, unbounded=False, "
                "neurons=128, layers=2, geo_feat_dim=15, n_levels<=16, num_semantic_classes<=32 "
                f"(got neurons={neurons}, layers={layers}, geo_feat_dim={geo_feat_dim}, n_levels={n_levels}, "
                f"num_semantic_classes={num_semantic_classes}, unbounded={unbounded})")

        self._meta, self._n_entries = hashgrid_levels(n_levels, base_resolution, max_resolution, log2_hashmap_size)
        enc_dim = 64  # the kernel's A tile is 64 wide; levels beyond n_levels read as zeros
        self._base_dims = [(neurons, enc_dim)] + [(neurons, neurons)] * (layers - 1) + [(16, neurons)]
        self._head_dims = [(neurons // 2, 32), (neurons // 2, neurons // 2), (16, neurons // 2)]
        self._sem_dims = [(neurons // 2, 16), (neurons // 2, neurons // 2), (32, neurons // 2)]  # kernel image
        # tcnn pads the output width to a multiple of 16: the flat ``mlp_sem.params`` has 16 output rows for C <= 16
        self._sem_dims_flat = self._sem_dims[:2] + [(_pad16(max(1, num_semantic_classes)), neurons // 2)]
        n_base_w = sum(a * b for a, b in self._base_dims)

        self.direction_encoding = _FlatParams(0, n_output_dims=16)
        self.mlp_base = _FlatParams(n_base_w + self._n_entries * 4, n_output_dims=1 + geo_feat_dim)
        self.mlp_head = _FlatParams(sum(a * b for a, b in self._head_dims), n_output_dims=3)
        if num_semantic_classes > 0:
            self.mlp_sem = _FlatParams(sum(a * b for a, b in self._sem_dims_flat), n_output_dims=num_semantic_classes)
        self._n_base_w = n_base_w
        self._cache = None
        self._cache_key = None
        self.reset_parameters()

    # -- initialisation: tcnn defaults (Xavier-uniform MLPs, grid U(-1e-4, 1e-4)) --
    @torch.no_grad()
    def reset_parameters(self, grid_range: float = 1e-4, generator=None):
        o = _xavier_(self.mlp_base.params, self._base_dims, generator)
        self.mlp_base.params[o:].uniform_(-grid_range, grid_range, generator=generator)
        _xavier_(self.mlp_head.params, self._head_dims, generator)
        if self.num_semantic_classes > 0:
            _xavier_(self.mlp_sem.params, self._sem_dims_flat, generator)

    # -- fp16 inference image of the parameters (rebuilt when a parameter changes) --
    def _packed(self):
        ps = [self.mlp_base.params, self.mlp_head.params]
        if self.num_semantic_classes > 0:
            ps.append(self.mlp_sem.params)
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in ps)
        if self._cache is not None and key == self._cache_key:
            return self._cache
        with torch.no_grad():
            dev = self.mlp_base.params.device

            def mats(flat, dims, rows=None):
                out, o = [], 0
                for i, (n_out, n_in) in enumerate(dims):
                    w = flat[o:o + n_out * n_in].reshape(n_out, n_in).to(torch.float16)
                    if rows is not None and rows[i] > n_out:  # zero rows up to the kernel's padded output width
                        w = torch.cat([w, w.new_zeros((rows[i] - n_out, n_in))])
                    out.append(_umma_pack(w))
                    o += n_out * n_in
                return out

            blobs = mats(self.mlp_base.params, self._base_dims) + mats(self.mlp_head.params, self._head_dims)
            if self.num_semantic_classes > 0:
                blobs += mats(self.mlp_sem.params, self._sem_dims_flat, rows=[a for a, _ in self._sem_dims])
            else:
                blobs += [torch.zeros(a * b, dtype=torch.float16, device=dev) for a, b in self._sem_dims]
            weights = torch.cat(blobs).contiguous()
            assert weights.numel() * 2 == int(LIB.raw("apnerf_field_weight_bytes")())
            table = self.mlp_base.params[self._n_base_w:].to(torch.float16).contiguous()
            self._cache = (weights, table)
            self._cache_key = key
        return self._cache

    def aabb_host(self) -> np.ndarray:
        """The aabb as a host float32[6] array for the C-ABI calls, read back from the device once (the buffer is
        normally constant; the copy is keyed on the buffer's storage and version counter, so ``load_state_dict`` or a
        move to another device refresh it)."""
        key = (self.aabb.data_ptr(), self.aabb._version, str(self.aabb.device))
        if getattr(self, "_aabb_host", None) is None or getattr(self, "_aabb_host_key", None) != key:
            self._aabb_host = np.ascontiguousarray(self.aabb.detach().cpu().numpy(), dtype=np.float32)
            self._aabb_host_key = key
        return self._aabb_host

    def _packed_t(self):
        """The weight blob of ``_packed`` with every matrix transposed ([in, out], same UMMA layout and
        offsets): the B operands of the backward kernel's dX = dY . W products."""
        ps = [self.mlp_base.params, self.mlp_head.params] + ([self.mlp_sem.params] if self.num_semantic_classes > 0 else [])
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in ps)
        if getattr(self, "_cache_t", None) is not None and key == self._cache_t_key:
            return self._cache_t
        with torch.no_grad():
            dev = self.mlp_base.params.device

            def mats(flat, dims, rows=None):
                out, o = [], 0
                for i, (n_out, n_in) in enumerate(dims):
                    w = flat[o:o + n_out * n_in].reshape(n_out, n_in).to(torch.float16)
                    if rows is not None and rows[i] > n_out:
                        w = torch.cat([w, w.new_zeros((rows[i] - n_out, n_in))])
                    out.append(_umma_pack(w.t().contiguous()))
                    o += n_out * n_in
                return out

            blobs = mats(self.mlp_base.params, self._base_dims) + mats(self.mlp_head.params, self._head_dims)
            if self.num_semantic_classes > 0:
                blobs += mats(self.mlp_sem.params, self._sem_dims_flat, rows=[a for a, _ in self._sem_dims])
            else:
                blobs += [torch.zeros(a * b, dtype=torch.float16, device=dev) for a, b in self._sem_dims]
            self._cache_t = torch.cat(blobs).contiguous()
            self._cache_t_key = key
        return self._cache_t

    def _run(self, positions, directions, density_only, return_feat=False):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._run_with_grad(positions, directions, density_only, return_feat)
        require_cuda(positions, directions, self.mlp_base.params)
        weights, table = self._packed()
        pos = positions.reshape(-1, 3).to(torch.float32).contiguous()
        n = pos.shape[0]
        dev = pos.device
        density = torch.empty(n, device=dev, dtype=torch.float32)
        aabb_host = self.aabb_host()
        feat = torch.empty((n, 15), device=dev, dtype=torch.float16) if return_feat else None
        rgb = sem = dirs = None
        C = self.num_semantic_classes
        if not density_only:
            dirs = directions.reshape(-1, 3).to(torch.float32).contiguous()
            rgb = torch.empty((n, 3), device=dev, dtype=torch.float32)
            sem = torch.empty((n, C), device=dev, dtype=torch.float32) if C > 0 else None
        if n:
            import ctypes
            with torch.cuda.device(dev):
                call("apnerf_field_forward", n, None, pos, dirs, None, None, None, None, None,
                     aabb_host.ctypes.data_as(ctypes.c_void_p), self.n_levels,
                     self._meta.ctypes.data_as(ctypes.c_void_p), table, weights, density,
                     rgb, 3, 1, sem, C, 1, C, feat, None, 1 if density_only else 0, 0)
        return density, rgb, sem, feat

    def _split(self, flat, dims):
        out, o = [], 0
        for n_out, n_in in dims:
            out.append(flat[o:o + n_out * n_in].view(n_out, n_in))
            o += n_out * n_in
        return out

    def _run_with_grad(self, positions, directions, density_only, return_feat):
        """Differentiable path (training): ``_FusedMLPs`` (two tcgen05 kernels + the hash-grid scatter) gives
        the raw network outputs; the activations of ngp.py:191-220 (trunc_exp, selector, sigmoid) are
        elementwise torch ops on top, so autograd reaches the flat fp32 ``params``."""
        require_cuda(positions, directions, self.mlp_base.params)
        if return_feat:
            raise NotImplementedError("query_density(return_feat=True) under autograd is not on the pipeline's path")
        pos = positions.reshape(-1, 3).to(torch.float32).contiguous().detach()
        if directions is None:  # query_density under autograd: the colour / semantic heads see a dummy direction
            dirs = torch.zeros_like(pos)
            dirs[:, 2] = 1.0
        else:
            dirs = directions.reshape(-1, 3).to(torch.float32).contiguous().detach()
        aabb_min, aabb_max = torch.split(self.aabb, 3, dim=-1)
        x = (pos - aabb_min) / (aabb_max - aabb_min)
        selector = ((x > 0.0) & (x < 1.0)).all(dim=-1)
        sem_params = self.mlp_sem.params if self.num_semantic_classes > 0 else None
        out = _FusedMLPs.apply(pos, dirs, self.mlp_base.params, self.mlp_head.params, sem_params,
                               self.direction_encoding.params, self)
        density = self.density_activation(out[0][:, None]) * selector[:, None]
        if density_only:
            return density.reshape(-1), None, None, None
        rgb = torch.sigmoid(out[1])
        sem = out[2] if self.num_semantic_classes > 0 else None
        return density.reshape(-1), rgb, sem, None

    def _apply(self, fn, *a, **k):
        self._cache = None
        self._cache_t = None
        self._aabb_host = None
        return super()._apply(fn, *a, **k)

    def query_density(self, x, return_feat: bool = False):
        """ngp.py:171-200: density [..., 1] (and the 15 geo features, upcast like ``.to(x)``)."""
        density, _, _, feat = self._run(x, None, True, return_feat)
        density = density.reshape(list(x.shape[:-1]) + [1])
        if return_feat:
            return density, feat.to(x.dtype).reshape(list(x.shape[:-1]) + [self.geo_feat_dim])
        return density

    def forward(self, positions: torch.Tensor, directions: torch.Tensor = None):
        """ngp.py:222-238: (rgb, density[, sem_logits])."""
        assert self.use_viewdirs and directions is not None
        assert positions.shape == directions.shape, f"{positions.shape} v.s. {directions.shape}"
        density, rgb, sem, _ = self._run(positions, directions, False)
        lead = list(positions.shape[:-1])
        rgb = rgb.reshape(lead + [3])
        density = density.reshape(lead + [1])
        if self.num_semantic_classes > 0:
            return rgb, density, sem.reshape(lead + [self.num_semantic_classes])
        return rgb, density
