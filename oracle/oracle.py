"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the render + score hot path.

numpy / ctypes front-end over ``apnerf_oracle.c`` plus numpy restatements of the pieces that
are plain array arithmetic.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product package never does (it fails loudly when its CUDA library is missing).

Every function cites the reference file:line it restates (paths relative to
/root/reference).

Parity status
-------------
* ray_aabb_intersect / traverse_grids: PINNED on the GPU box against the reference's own CUDA
  kernels (``oracle/_ref``; tests/test_gpu_reference.py), bit-exact.
* volume rendering (weights / accumulate): PINNED against the golden vectors of
  perception/nerfacc/tests/test_rendering.py:110-193, test_pack.py:10-17 and the docstring
  examples volrend.py:249-256,350-358, scan.py:77-80 (tests/test_oracle.py), and against the
  reference's own python (volrend.py imported on CPU) by tests/golden/make_golden.py.
* hash-grid / SH / MLP (tiny-cuda-nn): **parity unpinned** -- external unpinned dependency,
  absent from /root/reference (perception/models/requirements.txt:1); restated from its
  published algorithm.
* test-mode renderer (utils.py:782-1032), ray generation / subsample (habitat_to_data.py:274-301,
  462-467) and scoring (scripts/pipeline.py:666-798): PINNED against the reference's own Python,
  imported unmodified from /root/reference and run on CPU by tests/golden/make_golden.py (native
  calls served by apnerf_oracle.c, field by the restatement below); fixtures in
  tests/golden/reference_python_path.npz, checked by tests/test_golden.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libapnerf_oracle.so")
_lib = None
N_THREADS = int(os.environ.get("APNERF_ORACLE_THREADS", os.cpu_count() or 1))


def build():
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(HERE, "apnerf_oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _parallel(n, fn, grain=1024):
    """Run fn(i0, i1) over [0, n) on host threads (the C calls drop the GIL)."""
    if n <= 0:
        return
    nt = max(1, min(N_THREADS, (n + grain - 1) // grain))
    if nt == 1:
        fn(0, n)
        return
    chunk = (n + nt * 4 - 1) // (nt * 4)
    ranges = [(i, min(n, i + chunk)) for i in range(0, n, chunk)]
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(lambda r: fn(*r), ranges))


# --------------------------------------------------------------------------------------
# nerfacc/grid.py:13-51 -> csrc/grid.cu:284-313, 477-519
# --------------------------------------------------------------------------------------
def ray_aabb_intersect(rays_o, rays_d, aabbs, near_plane=-np.inf, far_plane=np.inf, miss_value=np.inf):
    rays_o, rays_d, aabbs = _f32(rays_o), _f32(rays_d), _f32(aabbs)
    n, m = rays_o.shape[0], aabbs.shape[0]
    t_mins = np.empty((n, m), np.float32)
    t_maxs = np.empty((n, m), np.float32)
    hits = np.empty((n, m), np.uint8)
    L = lib()

    def run(r0, r1):
        L.apo_ray_aabb_intersect(
            ctypes.c_int32(r0), ctypes.c_int32(r1), _p(rays_o), _p(rays_d),
            ctypes.c_float(near_plane), ctypes.c_float(far_plane), ctypes.c_int32(m), _p(aabbs),
            ctypes.c_float(miss_value), _p(t_mins), _p(t_maxs), _p(hits))

    _parallel(n, run)
    return t_mins, t_maxs, hits.astype(bool)


def _pure_ray_aabb_intersect(rays_o, rays_d, aabbs, near_plane=-np.inf, far_plane=np.inf, miss_value=np.inf):
    """nerfacc/grid.py:54-90 (the reference's own pure-torch oracle), in numpy."""
    rays_o, rays_d, aabbs = _f32(rays_o), _f32(rays_d), _f32(aabbs)
    amin, amax = aabbs[:, :3], aabbs[:, 3:]
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (amin[None] - rays_o[:, None]) / rays_d[:, None]
        t2 = (amax[None] - rays_o[:, None]) / rays_d[:, None]
    t_mins = np.minimum(t1, t2).max(-1)
    t_maxs = np.maximum(t1, t2).min(-1)
    hits = (t_maxs > t_mins) & (t_maxs > 0)
    t_mins = np.where(hits, np.clip(t_mins, near_plane, far_plane), miss_value).astype(np.float32)
    t_maxs = np.where(hits, np.clip(t_maxs, near_plane, far_plane), miss_value).astype(np.float32)
    return t_mins, t_maxs, hits


# --------------------------------------------------------------------------------------
# nerfacc/grid.py:93-192 + host code csrc/grid.cu:320-474 + data_spec.hpp:86-106
# --------------------------------------------------------------------------------------
def _run_traverse(rays_o, rays_d, rays_mask, res, binaries, aabbs, hits, t_sorted, t_indices, near, far,
                  step, cone, limit, first_pass, iv, sm, term):
    L = lib()
    n = rays_o.shape[0]
    n_grids = aabbs.shape[0]

    def run(r0, r1):
        L.apo_traverse_grids(
            ctypes.c_int32(r0), ctypes.c_int32(r1), _p(rays_o), _p(rays_d), _p(rays_mask),
            ctypes.c_int32(n_grids), _p(res), _p(binaries), _p(aabbs), _p(hits), _p(t_sorted),
            _p(t_indices), _p(near), _p(far), ctypes.c_float(step), ctypes.c_float(cone),
            ctypes.c_int32(limit), ctypes.c_int32(1 if first_pass else 0),
            _p(iv.get("vals")), _p(iv.get("ray_indices")), _p(iv.get("is_left")), _p(iv.get("is_right")),
            _p(iv.get("chunk_starts")), _p(iv.get("chunk_cnts")),
            _p(sm.get("vals")), _p(sm.get("ray_indices")), _p(sm.get("is_valid")),
            _p(sm.get("chunk_starts")), _p(sm.get("chunk_cnts")), _p(term))

    _parallel(n, run, grain=256)


def _alloc_from_chunk(spec, masks, valid):
    """data_spec.hpp:86-96 memalloc_data_from_chunk (zero-initialised)."""
    cnts = spec["chunk_cnts"]
    cum = np.cumsum(cnts, dtype=np.int64)
    n_edges = int(cum[-1]) if cum.size else 0
    spec["chunk_starts"] = (cum - cnts).astype(np.int64)
    spec["vals"] = np.zeros(n_edges, np.float32)
    spec["ray_indices"] = np.zeros(n_edges, np.int64)
    if masks:
        spec["is_left"] = np.zeros(n_edges, np.uint8)
        spec["is_right"] = np.zeros(n_edges, np.uint8)
    if valid:
        spec["is_valid"] = np.zeros(n_edges, np.uint8)


def traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=None, far_planes=None, step_size=1e-3,
                   cone_angle=0.0, traverse_steps_limit=None, over_allocate=False, rays_mask=None,
                   t_sorted=None, t_indices=None, hits=None):
    """Returns (intervals, samples, terminate_planes); intervals/samples are dicts with the
    RaySegmentsSpec fields (vals, ray_indices, is_left/is_right | is_valid, chunk_starts,
    chunk_cnts) plus ``packed_info``."""
    rays_o, rays_d, aabbs = _f32(rays_o), _f32(rays_d), _f32(aabbs)
    n = rays_o.shape[0]
    binaries = np.ascontiguousarray(binaries).astype(np.uint8)
    n_grids = binaries.shape[0]
    res = np.asarray(binaries.shape[1:], np.int32)
    near = np.zeros(n, np.float32) if near_planes is None else _f32(near_planes)
    far = np.full(n, np.inf, np.float32) if far_planes is None else _f32(far_planes)
    mask = np.ones(n, np.uint8) if rays_mask is None else np.ascontiguousarray(rays_mask).astype(np.uint8)
    limit = -1 if traverse_steps_limit is None else int(traverse_steps_limit)
    if over_allocate:
        assert limit > 0
    if t_sorted is None or t_indices is None or hits is None:
        t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, aabbs)
        cat = np.concatenate([t_mins, t_maxs], -1)
        t_indices = np.argsort(cat, axis=-1, kind="stable").astype(np.int64)
        t_sorted = np.take_along_axis(cat, t_indices, -1)
    t_sorted = _f32(t_sorted)
    t_indices = np.ascontiguousarray(t_indices, dtype=np.int64)
    hits = np.ascontiguousarray(hits).astype(np.uint8)
    term = np.zeros(n, np.float32)  # reference: torch.empty (masked rays keep garbage)
    iv, sm = {}, {}
    if over_allocate:  # grid.cu:364-404
        iv["chunk_cnts"] = (np.full(n, limit * 2, np.int64) * mask).astype(np.int64)
        _alloc_from_chunk(iv, True, False)
        sm["chunk_cnts"] = (np.full(n, limit, np.int64) * mask).astype(np.int64)
        _alloc_from_chunk(sm, False, True)
        _run_traverse(rays_o, rays_d, mask, res, binaries, aabbs, hits, t_sorted, t_indices, near, far,
                      step_size, cone_angle, limit, False, iv, sm, term)
        for s in (iv, sm):  # compute_chunk_start(): starts describe the COMPACTED order
            cum = np.cumsum(s["chunk_cnts"], dtype=np.int64)
            s["chunk_starts"] = (cum - s["chunk_cnts"]).astype(np.int64)
    else:  # grid.cu:405-470 two passes; rays_mask ignored (nullptr)
        iv["chunk_cnts"] = np.zeros(n, np.int64)
        sm["chunk_cnts"] = np.zeros(n, np.int64)
        _run_traverse(rays_o, rays_d, None, res, binaries, aabbs, hits, t_sorted, t_indices, near, far,
                      step_size, cone_angle, limit, True, iv, sm, None)
        _alloc_from_chunk(iv, True, False)
        _alloc_from_chunk(sm, False, True)
        _run_traverse(rays_o, rays_d, None, res, binaries, aabbs, hits, t_sorted, t_indices, near, far,
                      step_size, cone_angle, limit, False, iv, sm, term)
    for s in (iv, sm):
        s["packed_info"] = np.stack([s["chunk_starts"], s["chunk_cnts"]], -1)
        for k in ("is_left", "is_right", "is_valid"):
            if k in s:
                s[k] = s[k].astype(bool)
    return iv, sm, term


# --------------------------------------------------------------------------------------
# nerfacc/pack.py:10-49, nerfacc/scan.py:57-97, nerfacc/volrend.py:212-267, 315-365, 486-576
# --------------------------------------------------------------------------------------
def pack_info(ray_indices, n_rays):
    cnts = np.bincount(np.asarray(ray_indices, np.int64), minlength=n_rays).astype(np.int64)
    starts = np.cumsum(cnts) - cnts
    return np.stack([starts, cnts], -1)


def exclusive_sum(inputs, packed_info, backward=False):
    inputs = _f32(inputs)
    out = np.empty_like(inputs)
    starts = np.ascontiguousarray(packed_info[:, 0], dtype=np.int64)
    cnts = np.ascontiguousarray(packed_info[:, 1], dtype=np.int64)
    L = lib()

    def run(r0, r1):
        L.apo_exclusive_sum(ctypes.c_int32(r0), ctypes.c_int32(r1), _p(starts), _p(cnts), _p(inputs), _p(out),
                            ctypes.c_int32(1 if backward else 0))

    _parallel(starts.shape[0], run)
    return out


def render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None,
                                      n_rays=None, prefix_trans=None):
    if ray_indices is not None and packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    sigmas_dt = _f32(sigmas) * (_f32(t_ends) - _f32(t_starts))
    alphas = np.float32(1.0) - np.exp(-sigmas_dt)
    trans = np.exp(-exclusive_sum(sigmas_dt, packed_info))
    if prefix_trans is not None:
        trans = trans * _f32(prefix_trans)
    return trans.astype(np.float32), alphas.astype(np.float32)


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                               prefix_trans=None):
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices,
                                                      n_rays, prefix_trans)
    return (trans * alphas).astype(np.float32), trans, alphas


def render_visibility_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                                   early_stop_eps=1e-4, alpha_thre=0.0, prefix_trans=None):
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices,
                                                      n_rays, prefix_trans)
    vis = trans >= np.float32(early_stop_eps)
    if alpha_thre > 0:
        vis = vis & (alphas >= np.float32(alpha_thre))
    return vis


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    weights = _f32(weights)
    src = weights[:, None] if values is None else weights[:, None] * _f32(values)
    out = np.zeros((n_rays, src.shape[-1]), np.float32)
    np.add.at(out, np.asarray(ray_indices, np.int64), src)
    return out


def accumulate_along_rays_(weights, values, ray_indices, outputs):
    weights = _f32(weights)
    src = weights[:, None] if values is None else weights[:, None] * _f32(values)
    np.add.at(outputs, np.asarray(ray_indices, np.int64), src.astype(np.float32))


# --------------------------------------------------------------------------------------
# tiny-cuda-nn restatement -- parity unpinned
# --------------------------------------------------------------------------------------
def hashgrid_meta(n_levels=16, base_resolution=16, max_resolution=4096, log2_hashmap_size=19):
    """Per-level {scale (f32 bits), resolution, size, offset} in table entries
    (perception/models/radiance_fields/ngp.py:103-105,123-133)."""
    s = np.exp((np.log(max_resolution) - np.log(base_resolution)) / (n_levels - 1))
    meta = np.zeros((n_levels, 4), np.uint32)
    off = 0
    for l in range(n_levels):
        scale = np.exp2(l * np.log2(s)) * base_resolution - 1.0
        res = int(np.ceil(scale)) + 1
        size = min((res ** 3 + 7) // 8 * 8, 1 << log2_hashmap_size)
        meta[l] = (np.float32(scale).view(np.uint32), res, size, off)
        off += size
    return meta, off


def hashgrid_encode(x01, table_half, meta, want_indices=False, parallel=True):
    """x01 [N,3] fp32 -> fp16 features [N, L*4] (+ table entry indices [N,L,8])."""
    x01 = _f32(x01)
    n = x01.shape[0]
    L_ = meta.shape[0]
    table = np.ascontiguousarray(table_half).view(np.uint16)
    out = np.empty((n, L_ * 4), np.uint16)
    idx = np.empty((n, L_, 8), np.uint32) if want_indices else None
    meta = np.ascontiguousarray(meta, dtype=np.uint32)
    L = lib()

    def run(s0, s1):
        L.apo_hashgrid_encode(ctypes.c_int64(s0), ctypes.c_int64(s1), _p(x01), ctypes.c_int32(L_), _p(meta),
                              _p(table), _p(out), _p(idx))

    if parallel:
        _parallel(n, run, grain=256)
    elif n:
        run(0, n)
    enc = out.view(np.float16)
    return (enc, idx) if want_indices else enc


def sh4(dirs, parallel=True):
    dirs = _f32(dirs)
    n = dirs.shape[0]
    out = np.empty((n, 16), np.uint16)
    L = lib()
    run = lambda s0, s1: L.apo_sh4(ctypes.c_int64(s0), ctypes.c_int64(s1), _p(dirs), _p(out))
    if parallel:
        _parallel(n, run)
    elif n:
        run(0, n)
    return out.view(np.float16)


def mlp_forward(x_half, weights_half):
    """FullyFusedMLP restatement: fp16 operands, fp32 accumulate, ReLU on hidden layers, each
    layer's output rounded to fp16 (weights row-major [out, in], no bias)."""
    h = np.asarray(x_half, np.float16)
    for i, w in enumerate(weights_half):
        y = h.astype(np.float32) @ np.asarray(w, np.float16).astype(np.float32).T
        if i < len(weights_half) - 1:
            y = np.maximum(y, 0.0)
        h = y.astype(np.float16)
    return h


class FieldParams:
    """Splits the three flat fp32 ``params`` vectors (tcnn layout: MLP weights first, then the
    grid table) into fp16 matrices.  ngp.py:123-169."""

    def __init__(self, base_params, head_params, sem_params, neurons=128, n_hidden=2, geo_feat_dim=15,
                 num_semantic_classes=29, n_levels=16, base_resolution=16, max_resolution=4096,
                 log2_hashmap_size=19):
        self.meta, self.n_entries = hashgrid_meta(n_levels, base_resolution, max_resolution, log2_hashmap_size)
        self.num_semantic_classes = num_semantic_classes
        self.geo_feat_dim = geo_feat_dim
        enc_dim = n_levels * 4
        out_dim = _pad16(1 + geo_feat_dim)

        def split(flat, dims):
            ws, o = [], 0
            for (n_out, n_in) in dims:
                ws.append(np.asarray(flat[o:o + n_out * n_in], np.float32).reshape(n_out, n_in).astype(np.float16))
                o += n_out * n_in
            return ws, o

        dims = [(neurons, enc_dim)] + [(neurons, neurons)] * (n_hidden - 1) + [(out_dim, neurons)]
        self.base_w, o = split(base_params, dims)
        self.table = np.asarray(base_params[o:o + self.n_entries * 4], np.float32).astype(np.float16).reshape(-1, 4)
        hn = neurons // 2
        self.head_in = _pad16(16 + geo_feat_dim)
        self.head_w, _ = split(head_params, [(hn, self.head_in), (hn, hn), (_pad16(3), hn)])
        self.sem_in = _pad16(geo_feat_dim)
        if num_semantic_classes > 0:
            self.sem_w, _ = split(sem_params, [(hn, self.sem_in), (hn, hn), (_pad16(num_semantic_classes), hn)])


def _pad16(n):
    return (n + 15) // 16 * 16


def field_forward(positions, directions, aabb, fp: FieldParams, density_only=False, chunk=8192):
    """NGPRadianceField.forward / query_density, perception/models/radiance_fields/ngp.py:171-238.
    Rows are independent, so large batches are evaluated in row chunks on host threads (numpy's fp16 <-> fp32
    conversions are single-threaded and were 55 % of the oracle's time); per-row results do not depend on the
    chunking (checked by tests/test_oracle.py)."""
    n = np.asarray(positions).shape[0]
    if n <= chunk or N_THREADS == 1:
        return _field_forward_rows(positions, directions, aabb, fp, density_only)
    positions = _f32(positions)
    directions = None if directions is None else _f32(directions)
    ranges = [(i, min(n, i + chunk)) for i in range(0, n, chunk)]
    with _blas_threads(1), ThreadPoolExecutor(min(N_THREADS, len(ranges))) as ex:
        parts = list(ex.map(lambda r: _field_forward_rows(positions[r[0]:r[1]],
                                                          None if directions is None else directions[r[0]:r[1]],
                                                          aabb, fp, density_only, parallel=False), ranges))
    if density_only:
        return np.concatenate(parts, 0)
    return tuple(np.concatenate([p[k] for p in parts], 0) for k in range(len(parts[0])))


class _blas_threads:
    """Limit the BLAS pool while our own threads run one GEMM each (threadpoolctl when present)."""

    def __init__(self, n):
        self.n, self.ctx = n, None

    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits

            self.ctx = threadpool_limits(limits=self.n, user_api="blas")
            self.ctx.__enter__()
        except Exception:
            self.ctx = None
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _field_forward_rows(positions, directions, aabb, fp: FieldParams, density_only=False, parallel=True):
    positions = _f32(positions)
    aabb = _f32(aabb)
    x = (positions - aabb[:3]) / (aabb[3:] - aabb[:3])
    selector = ((x > 0.0) & (x < 1.0)).all(-1)
    enc = hashgrid_encode(x, fp.table, fp.meta, parallel=parallel)
    base = mlp_forward(enc, fp.base_w).astype(np.float32)
    density = np.exp(base[:, :1] - np.float32(1.0)) * selector[:, None].astype(np.float32)
    if density_only:
        return density.astype(np.float32)
    feat = base[:, 1:1 + fp.geo_feat_dim].astype(np.float16)
    n = positions.shape[0]
    h = np.ones((n, fp.head_in), np.float16)  # tcnn pads inputs to a multiple of 16 with 1.0
    h[:, :16] = sh4(directions, parallel=parallel)
    h[:, 16:16 + fp.geo_feat_dim] = feat
    rgb_raw = mlp_forward(h, fp.head_w).astype(np.float32)[:, :3]
    rgb = (1.0 / (1.0 + np.exp(-rgb_raw))).astype(np.float32)
    if fp.num_semantic_classes > 0:
        s = np.ones((n, fp.sem_in), np.float16)
        s[:, :fp.geo_feat_dim] = feat
        sem = mlp_forward(s, fp.sem_w).astype(np.float32)[:, :fp.num_semantic_classes]
        return rgb, density.astype(np.float32), sem
    return rgb, density.astype(np.float32)


# --------------------------------------------------------------------------------------
# perception/data_proc/habitat_to_data.py:274-301 generate_image_rays, :461-467 subsample
# --------------------------------------------------------------------------------------
def generate_image_rays(pose, width, height, focal):
    """pose [4,4] fp32 camera-to-world (OpenGL); returns origins, viewdirs [H*W, 3] fp32."""
    pose = _f32(pose)
    x, y = np.meshgrid(np.arange(width), np.arange(height), indexing="xy")
    x = x.flatten().astype(np.float32)
    y = y.flatten().astype(np.float32)
    cx, cy = np.float32(width / 2), np.float32(height / 2)
    f = np.float32(focal)
    cam = np.stack([(x - cx + np.float32(0.5)) / f, (y - cy + np.float32(0.5)) / f * np.float32(-1.0),
                    np.full_like(x, -1.0)], 1)
    prod = cam[:, None, :] * pose[None, :3, :3]
    d = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    o = np.broadcast_to(pose[:3, 3], d.shape).copy()
    nrm = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    return o, (d / nrm[:, None]).astype(np.float32)


def subsample_indices(n_total, n_keep):
    return np.round(np.linspace(0, n_total - 1, n_keep)).astype(int)


# --------------------------------------------------------------------------------------
# perception/models/utils.py:782-1032 render_probablistic_image_with_occgrid_test
# --------------------------------------------------------------------------------------
def render_probablistic_image_with_occgrid_test(max_samples, field_fn, binaries, aabbs, rays_o, rays_d,
                                                num_semantic_classes, near_plane=0.0, far_plane=1e10,
                                                render_step_size=1e-3, render_bkgd=None, cone_angle=0.0,
                                                alpha_thre=0.0, early_stop_eps=1e-4, trace=None, ray_counts=None):
    """field_fn(positions, dirs) -> (rgb [N,3], sigma [N,1], sem [N,C]).  Returns
    (rgb, rgb_var, opacity, depth, depth_var, sem, total_samples).  ``trace`` (a list) receives
    one dict per marching iteration (n_alive, n_samples, ray_indices, t_starts, t_ends) so the
    fused GPU path's device-side schedule can be compared step by step.  ``ray_counts`` (a dict)
    receives per-ray int64 totals ``evaluated`` (samples marched and sent through the field) and
    ``visible`` (samples that passed alpha_thre and were composited): two renders took the same
    discrete decisions for a ray iff both totals agree (the parity tests' definition of a
    threshold-flip ray)."""
    rays_o, rays_d = _f32(rays_o), _f32(rays_d)
    num_rays = rays_o.shape[0]
    C = num_semantic_classes
    opacity = np.zeros((num_rays, 1), np.float32)
    depth = np.zeros((num_rays, 1), np.float32)
    rgb = np.zeros((num_rays, 3), np.float32)
    sem = np.zeros((num_rays, C), np.float32)
    depth_var = np.zeros((num_rays, 1), np.float32)
    rgb_var = np.zeros((num_rays, 3), np.float32)
    ray_mask = np.ones(num_rays, bool)
    min_samples = 1 if cone_angle == 0 else 4
    iter_samples = total_samples = 0
    near_planes = np.full(num_rays, near_plane, np.float32)
    far_planes = np.full(num_rays, far_plane, np.float32)
    t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, aabbs)
    n_grids = binaries.shape[0]
    cat = np.concatenate([t_mins, t_maxs], -1)
    if n_grids > 1:
        t_indices = np.argsort(cat, -1, kind="stable").astype(np.int64)
        t_sorted = np.take_along_axis(cat, t_indices, -1)
    else:
        t_sorted = cat
        t_indices = np.broadcast_to(np.arange(2 * n_grids, dtype=np.int64), (num_rays, 2 * n_grids)).copy()
    opc_thre = np.float32(1 - early_stop_eps)
    if render_bkgd is None:
        render_bkgd = np.zeros(3, np.float32)
    while iter_samples < max_samples:
        n_alive = int(ray_mask.sum())
        if n_alive == 0:
            break
        n_samples = max(min(num_rays // n_alive, 64), min_samples)
        iter_samples += n_samples
        intervals, samples, termination_planes = traverse_grids(
            rays_o, rays_d, binaries, aabbs, near_planes, far_planes, render_step_size, cone_angle,
            n_samples, True, ray_mask, t_sorted, t_indices, hits)
        t_starts = intervals["vals"][intervals["is_left"]]
        t_ends = intervals["vals"][intervals["is_right"]]
        ray_indices = samples["ray_indices"][samples["is_valid"]]
        packed_info = samples["packed_info"]
        if trace is not None:
            trace.append(dict(n_alive=n_alive, n_samples=n_samples, ray_indices=ray_indices.copy(),
                              t_starts=t_starts.copy(), t_ends=t_ends.copy()))
        if ray_counts is not None:
            ray_counts.setdefault("evaluated", np.zeros(num_rays, np.int64))
            ray_counts.setdefault("visible", np.zeros(num_rays, np.int64))
            ray_counts["evaluated"] += np.bincount(ray_indices, minlength=num_rays)
        positions = rays_o[ray_indices] + rays_d[ray_indices] * (t_starts[:, None] + t_ends[:, None]) / np.float32(2.0)
        if positions.shape[0] == 0:
            rgbs = np.zeros((0, 3), np.float32)
            sigmas = np.zeros((0,), np.float32)
            sems = np.zeros((0, C), np.float32)
        else:
            rgbs, sigmas, sems = field_fn(positions, rays_d[ray_indices])
            sigmas = sigmas.reshape(-1)
        weights, _, alphas = render_weight_from_density(
            t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=num_rays,
            prefix_trans=1 - opacity[ray_indices, 0])
        if alpha_thre > 0:
            vis = alphas >= np.float32(alpha_thre)
            ray_indices, rgbs, weights, t_starts, t_ends, sems = (
                ray_indices[vis], rgbs[vis], weights[vis], t_starts[vis], t_ends[vis], sems[vis])
        if ray_counts is not None:
            ray_counts["visible"] += np.bincount(ray_indices, minlength=num_rays)
        t_mid = (t_starts + t_ends)[:, None] / np.float32(2.0)
        accumulate_along_rays_(weights, rgbs, ray_indices, rgb)
        accumulate_along_rays_(weights, None, ray_indices, opacity)
        accumulate_along_rays_(weights, t_mid, ray_indices, depth)
        accumulate_along_rays_(weights, sems, ray_indices, sem)
        accumulate_along_rays_(weights, (rgbs - rgb[ray_indices]) ** 2, ray_indices, rgb_var)
        accumulate_along_rays_(weights, (t_mid - depth[ray_indices]) ** 2, ray_indices, depth_var)
        near_planes = termination_planes
        ray_mask = (opacity.reshape(-1) <= opc_thre) & (packed_info[:, 1] == n_samples)
        total_samples += ray_indices.shape[0]
    rgb = rgb + np.asarray(render_bkgd, np.float32) * (1.0 - opacity)
    depth = depth / np.maximum(opacity, np.finfo(np.float32).eps)
    return rgb, rgb_var, opacity, depth, depth_var, sem, total_samples


# --------------------------------------------------------------------------------------
# scripts/pipeline.py:727-781 predictive information from the ensemble's renders
# --------------------------------------------------------------------------------------
def predictive_information(rgb_var, depth_var, acc, sem):
    """Inputs are float64 arrays stacked over the ensemble on axis 0:
    rgb_var [E,...,3], depth_var [E,...], acc [E,...], sem [E,...,C].
    Returns (rgb_pi, depth_pi, 3*sem_pi, 2*occ_pi) -- the four entries the reference appends
    to trajector_uncertainty_list (pipeline.py:783-790); their sum is the trajectory score."""
    rgb_var = np.asarray(rgb_var, np.float64)
    depth_var = np.asarray(depth_var, np.float64)
    acc = np.asarray(acc, np.float64)
    sem = np.asarray(sem, np.float64)
    c = 2 * np.pi * np.e
    rgb_ce = np.log(c * rgb_var + 1e-4) / 2
    rgb_pi = np.mean(np.log(c * (np.sum(rgb_var, 0) / 2) + 1e-4) / 2 - np.mean(rgb_ce, 0))
    d_ce = np.log(c * depth_var + 1e-4) / 2
    depth_pi = np.mean(np.log(c * (np.sum(depth_var, 0) / 2) + 1e-4) / 2 - np.mean(d_ce, 0))
    m = sem.max(-1, keepdims=True)
    e = np.exp(sem - m)
    p = e / e.sum(-1, keepdims=True)
    sem_ce = -np.sum((p + 1e-4) * np.log(p + 1e-4), -1)
    pe = np.mean(p, 0)
    sem_pi = np.mean(-np.sum((pe + 1e-4) * np.log(pe + 1e-4), -1) - np.mean(sem_ce, 0))

    def bern(a):
        return -(a + 1e-4) * np.log(a + 1e-4) - (1 - a + 1e-4) * np.log(1 - a + 1e-4)

    occ_pi = np.mean(bern(np.mean(acc, 0)) - np.mean(bern(acc), 0))
    return np.array([rgb_pi, depth_pi, sem_pi * 3, occ_pi * 2], np.float64)
