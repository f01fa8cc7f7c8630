"""GPU: the CUDA path against the REFERENCE'S OWN kernels (oracle/_ref, built from
/root/reference/perception/nerfacc/nerfacc/cuda/csrc/{grid,scan}.cu) and against the C oracle.
Ray-march outputs are integer/index-like work: the bar is bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _scene(n_rays, n_grids, res, seed, inside=True):
    g = torch.Generator().manual_seed(seed)
    rays_o = (torch.rand((n_rays, 3), generator=g) * 2 - 1) * (0.9 if inside else 3.0)
    rays_d = torch.randn((n_rays, 3), generator=g)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    # a few axis-aligned and zero-component directions (inv_dir = +-inf paths)
    rays_d[0] = torch.tensor([1.0, 0.0, 0.0])
    rays_d[1] = torch.tensor([0.0, -1.0, 0.0])
    rays_d[2] = torch.tensor([0.6, 0.8, 0.0])
    base = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
    aabbs = []
    for i in range(n_grids):
        c, e = (base[:3] + base[3:]) / 2, (base[3:] - base[:3]) / 2 * 2 ** i
        aabbs.append(torch.cat([c - e, c + e]))
    aabbs = torch.stack(aabbs)
    binaries = torch.rand((n_grids, res, res, res), generator=g) > 0.7
    return rays_o.to(DEV), rays_d.to(DEV), binaries.to(DEV), aabbs.to(DEV)


def _ref_traverse(ref, rays_o, rays_d, binaries, aabbs, near, far, step, cone, limit, over, mask):
    t_mins, t_maxs, hits = ref.ray_aabb_intersect(rays_o, rays_d, aabbs, -float("inf"), float("inf"), float("inf"))
    t_sorted, t_indices = torch.sort(torch.cat([t_mins, t_maxs], -1), -1)
    iv, sm, term = ref.traverse_grids(rays_o, rays_d, mask, binaries, aabbs, t_sorted.contiguous(),
                                      t_indices.contiguous(), hits, near, far, step, cone, True, True, True,
                                      limit, over)
    return iv, sm, term, (t_sorted, t_indices, hits)


def _same(a, b, what):
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if a.dtype.is_floating_point:
        same = (a.view(torch.int32) == b.view(torch.int32)) if a.dtype == torch.float32 else (a == b)
    else:
        same = a == b
    assert bool(same.all()), f"{what}: {(~same).sum().item()} / {same.numel()} elements differ"


@pytest.mark.parametrize("n_grids,res,step,cone,near", [
    (1, 128, 1e-3, 0.004, 0.1),   # the pipeline's configuration (config_*.yaml:13,27-29)
    (1, 64, 5e-3, 0.0, 0.0),
    (4, 32, 1e-3, 0.0, 0.0),      # tests/test_grid.py:39-68
    (4, 32, 2e-3, 0.01, 0.05),
    (2, 16, 0.0, 0.0, 0.0),       # step_size <= 0: one sample per crossed cell
])
def test_traverse_train_mode_bit_exact(apnerf, ref_cuda, oracle, n_grids, res, step, cone, near):
    from apnerf import nerfacc

    n = 3000
    rays_o, rays_d, binaries, aabbs = _scene(n, n_grids, res, seed=7 + res)
    near_p = torch.full((n,), near, device=DEV)
    far_p = torch.full((n,), 1e10, device=DEV)
    mask = torch.ones(n, dtype=torch.bool, device=DEV)
    riv, rsm, rterm, isect = _ref_traverse(ref_cuda, rays_o, rays_d, binaries, aabbs, near_p, far_p, step, cone,
                                           -1, False, mask)
    iv, sm, term = nerfacc.traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=near_p, far_planes=far_p,
                                          step_size=step, cone_angle=cone)
    assert rsm.vals.numel() > 1000
    _same(sm.packed_info[:, 1], rsm.chunk_cnts, "sample counts")
    _same(sm.packed_info[:, 0], rsm.chunk_starts, "sample starts")
    _same(sm.ray_indices, rsm.ray_indices, "ray_indices")
    _same(sm.vals, rsm.vals, "sample t_mid")
    _same(iv.vals, riv.vals, "interval edges")
    _same(iv.is_left, riv.is_left, "is_left")
    _same(iv.is_right, riv.is_right, "is_right")
    _same(iv.ray_indices, riv.ray_indices, "interval ray_indices")
    _same(iv.packed_info[:, 1], riv.chunk_cnts, "interval counts")
    # rays without samples are skipped by the fill pass (grid.cu:103-106): their terminate plane is
    # uninitialised memory (torch::empty) in the reference, so only rays with samples are compared
    has = rsm.chunk_cnts > 0
    _same(term[has], rterm[has], "terminate planes")
    # ... and the C oracle agrees with the reference kernel too (this is what pins the oracle)
    t_sorted, t_indices, hits = (t.cpu().numpy() for t in isect)
    oiv, osm, oterm = oracle.traverse_grids(rays_o.cpu().numpy(), rays_d.cpu().numpy(), binaries.cpu().numpy(),
                                            aabbs.cpu().numpy(), near_p.cpu().numpy(), far_p.cpu().numpy(), step,
                                            cone, t_sorted=t_sorted, t_indices=t_indices, hits=hits)
    assert (osm["chunk_cnts"] == rsm.chunk_cnts.cpu().numpy()).all()
    assert (oiv["vals"].view(np.int32) == riv.vals.cpu().numpy().view(np.int32)).all()
    has = has.cpu().numpy()
    assert (oterm.view(np.int32)[has] == rterm.cpu().numpy().view(np.int32)[has]).all()


@pytest.mark.parametrize("limit,cone", [(4, 0.004), (64, 0.004), (7, 0.0)])
def test_traverse_test_mode_bit_exact(apnerf, ref_cuda, limit, cone):
    """over_allocate + rays_mask + per-ray near planes: three chained iterations like the
    test-mode renderer (perception/models/utils.py:896-1009)."""
    from apnerf import nerfacc

    n = 2000
    rays_o, rays_d, binaries, aabbs = _scene(n, 1, 128, seed=11)
    near_r = torch.full((n,), 0.1, device=DEV)
    near_m = near_r.clone()
    far_p = torch.full((n,), 1e10, device=DEV)
    mask_r = torch.rand(n, device=DEV) > 0.2
    mask_m = mask_r.clone()
    t_mins, t_maxs, hits = nerfacc.ray_aabb_intersect(rays_o, rays_d, aabbs)
    t_sorted = torch.cat([t_mins, t_maxs], -1)
    t_indices = torch.arange(2, device=DEV, dtype=torch.int64).expand(n, 2).contiguous()
    for it in range(3):
        riv, rsm, rterm = ref_cuda.traverse_grids(rays_o, rays_d, mask_r, binaries, aabbs, t_sorted, t_indices, hits,
                                                  near_r, far_p, 1e-3, cone, True, True, True, limit, True)
        iv, sm, term = nerfacc.traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=near_m, far_planes=far_p,
                                              step_size=1e-3, cone_angle=cone, traverse_steps_limit=limit,
                                              over_allocate=True, rays_mask=mask_m, t_sorted=t_sorted,
                                              t_indices=t_indices, hits=hits)
        _same(sm.packed_info[:, 1], rsm.chunk_cnts, f"it{it} counts")
        _same(sm.packed_info[:, 0], rsm.chunk_starts, f"it{it} starts")
        _same(iv.vals[iv.is_left], riv.vals[riv.is_left], f"it{it} t_starts")
        _same(iv.vals[iv.is_right], riv.vals[riv.is_right], f"it{it} t_ends")
        _same(sm.ray_indices[sm.is_valid], rsm.ray_indices[rsm.is_valid], f"it{it} ray_indices")
        _same(term[mask_m], rterm[mask_r], f"it{it} terminate planes")
        # masked rays keep an unspecified terminate plane in the reference (torch::empty); carry ours
        near_r = torch.where(mask_r, rterm, near_r)
        near_m = torch.where(mask_m, term, near_m)
        mask_r = mask_r & (rsm.chunk_cnts == limit)
        mask_m = mask_m & (sm.packed_info[:, 1] == limit)


def test_ray_aabb_bit_exact(apnerf, ref_cuda, oracle):
    from apnerf import nerfacc

    torch.manual_seed(42)
    rays_o = torch.rand((1000, 3), device=DEV)
    rays_d = torch.randn((1000, 3), device=DEV)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    amin = torch.rand((100, 3), device=DEV)
    aabbs = torch.cat([amin, amin + torch.rand_like(amin)], -1)
    for near, far in [(-float("inf"), float("inf")), (0.1, 1.5)]:
        r = ref_cuda.ray_aabb_intersect(rays_o, rays_d, aabbs, near, far, float("inf"))
        m = nerfacc.ray_aabb_intersect(rays_o, rays_d, aabbs, near, far)
        o = oracle.ray_aabb_intersect(rays_o.cpu().numpy(), rays_d.cpu().numpy(), aabbs.cpu().numpy(), near, far)
        for a, b, c, name in zip(m, r, o, ("t_mins", "t_maxs", "hits")):
            _same(a, b, name)
            assert (a.cpu().numpy() == c).all(), f"oracle {name}"


def test_exclusive_sum_vs_reference(apnerf, ref_cuda):
    """fp32 packed scan: summation order differs from the reference's Blelloch tree, so the
    bar is the compositing tolerance (<= 1e-5 relative) instead of bit equality."""
    from apnerf.nerfacc import exclusive_sum, inclusive_sum

    g = torch.Generator().manual_seed(0)
    cnts = torch.randint(0, 1025, (4096,), generator=g)
    starts = torch.cumsum(cnts, 0) - cnts
    n = int(cnts.sum())
    x = torch.rand(n, generator=g).to(DEV) * 1e-2
    starts, cnts = starts.to(DEV), cnts.to(DEV)
    pi = torch.stack([starts, cnts], -1)
    for backward in (False, True):
        ref = ref_cuda.exclusive_sum(starts, cnts, x, False, backward)
        from apnerf.nerfacc.scan import _packed_sum

        mine = _packed_sum(starts, cnts, x, False, False, backward)
        scale = ref.abs().max()
        assert ((mine - ref).abs() <= 1e-5 * torch.maximum(ref.abs(), scale * 1e-2)).all()
        refi = ref_cuda.inclusive_sum(starts, cnts, x, False, backward)
        minei = _packed_sum(starts, cnts, x, True, False, backward)
        assert ((minei - refi).abs() <= 1e-5 * torch.maximum(refi.abs(), scale * 1e-2)).all()
    assert exclusive_sum(x, pi).shape == x.shape and inclusive_sum(x, pi).shape == x.shape
