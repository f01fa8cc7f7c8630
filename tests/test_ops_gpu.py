"""GPU: the reference's own nerfacc test-suite cases (perception/nerfacc/tests/test_grid.py,
test_rendering.py, test_scan.py, test_pack.py) run against the drop-in ops, plus the config-1
integration sized check against the oracle (fp32 compositing <= 1e-5 relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
device = "cuda:0"


def _query(x, data, base_aabb):  # perception/nerfacc/nerfacc/grid.py:201-237
    aabb_min, aabb_max = torch.split(base_aabb, 3, dim=0)
    x_norm = (x - aabb_min) / (aabb_max - aabb_min)
    maxval = (x_norm - 0.5).abs().max(dim=-1).values
    maxval = torch.clamp(maxval, min=0.1)
    exponent = torch.frexp(maxval)[1].long()
    mip = torch.clamp(exponent + 1, min=0)
    selector = mip < data.shape[0]
    scale = 2 ** mip
    x_unit = (x_norm - 0.5) / scale[:, None] + 0.5
    resolution = torch.tensor(data.shape[1:], device=x.device)
    ix = (x_unit * resolution).long()
    ix = torch.clamp(ix, max=resolution - 1)
    mip = torch.clamp(mip, max=data.shape[0] - 1)
    return data[mip, ix[:, 0], ix[:, 1], ix[:, 2]] * selector, selector


def test_ray_aabb_intersect(apnerf):
    from apnerf.nerfacc import ray_aabb_intersect

    torch.manual_seed(42)
    rays_o = torch.rand((1000, 3), device=device)
    rays_d = torch.randn((1000, 3), device=device)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    aabb_min = torch.rand((100, 3), device=device)
    aabb_max = aabb_min + torch.rand_like(aabb_min)
    aabbs = torch.cat([aabb_min, aabb_max], dim=-1)
    tmins, tmaxs, hits = ray_aabb_intersect(rays_o, rays_d, aabbs)
    t1 = (aabb_min[None] - rays_o[:, None]) / rays_d[:, None]
    t2 = (aabb_max[None] - rays_o[:, None]) / rays_d[:, None]
    _tmins = torch.max(torch.min(t1, t2), dim=-1)[0]
    _tmaxs = torch.min(torch.max(t1, t2), dim=-1)[0]
    _hits = (_tmaxs > _tmins) & (_tmaxs > 0)
    inf = torch.tensor(float("inf"), device=device)
    assert torch.allclose(tmins, torch.where(_hits, _tmins, inf))
    assert torch.allclose(tmaxs, torch.where(_hits, _tmaxs, inf))
    assert (hits == _hits).all()


def test_traverse_grids_samples_in_occupied_cells(apnerf):
    from apnerf.nerfacc import traverse_grids
    from apnerf.nerfacc.grid import _enlarge_aabb

    torch.manual_seed(42)
    rays_o = torch.randn((10, 3), device=device)
    rays_d = torch.randn((10, 3), device=device)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    base_aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], device=device)
    aabbs = torch.stack([_enlarge_aabb(base_aabb, 2 ** i) for i in range(4)])
    binaries = torch.rand((4, 32, 32, 32), device=device) > 0.5
    intervals, samples, _ = traverse_grids(rays_o, rays_d, binaries, aabbs)
    ray_indices = samples.ray_indices
    t_starts = intervals.vals[intervals.is_left]
    t_ends = intervals.vals[intervals.is_right]
    positions = rays_o[ray_indices] + rays_d[ray_indices] * (t_starts + t_ends)[:, None] / 2.0
    occs, selector = _query(positions, binaries, base_aabb)
    assert occs.float().mean() > 0.9999, occs.float().mean()
    assert selector.all()


def test_traverse_grids_test_mode(apnerf):
    from apnerf.nerfacc import accumulate_along_rays, traverse_grids
    from apnerf.nerfacc.grid import _enlarge_aabb

    torch.manual_seed(42)
    n_rays = 10
    rays_o = torch.randn((n_rays, 3), device=device)
    rays_d = torch.randn((n_rays, 3), device=device)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    base_aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], device=device)
    aabbs = torch.stack([_enlarge_aabb(base_aabb, 2 ** i) for i in range(4)])
    binaries = torch.rand((4, 32, 32, 32), device=device) > 0.5
    intervals, samples, _ = traverse_grids(rays_o, rays_d, binaries, aabbs)
    ray_indices = samples.ray_indices
    t_starts = intervals.vals[intervals.is_left]
    t_ends = intervals.vals[intervals.is_right]
    accum_t_starts = accumulate_along_rays(t_starts, None, ray_indices, n_rays)
    accum_t_ends = accumulate_along_rays(t_ends, None, ray_indices, n_rays)
    _accum_t_starts, _accum_t_ends = 0.0, 0.0
    _terminate_planes, _rays_mask = None, None
    for _ in range(2):
        _intervals, _samples, _terminate_planes = traverse_grids(
            rays_o, rays_d, binaries, aabbs, near_planes=_terminate_planes, traverse_steps_limit=4000,
            over_allocate=True, rays_mask=_rays_mask)
        _rays_mask = _samples.packed_info[:, 1] == 4000
        _ray_indices = _samples.ray_indices[_samples.is_valid]
        _t_starts = _intervals.vals[_intervals.is_left]
        _t_ends = _intervals.vals[_intervals.is_right]
        _accum_t_starts += accumulate_along_rays(_t_starts, None, _ray_indices, n_rays)
        _accum_t_ends += accumulate_along_rays(_t_ends, None, _ray_indices, n_rays)
    assert (~_rays_mask).all()
    assert torch.allclose(_accum_t_starts, accum_t_starts, atol=1e-1)
    assert torch.allclose(accum_t_ends, _accum_t_ends, atol=1e-1)


def test_traverse_grids_with_near_far_planes(apnerf):
    from apnerf.nerfacc import traverse_grids

    rays_o = torch.tensor([[-1.0, 0.0, 0.0]], device=device)
    rays_d = torch.tensor([[1.0, 0.01, 0.01]], device=device)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    binaries = torch.ones((1, 1, 1, 1), dtype=torch.bool, device=device)
    aabbs = torch.tensor([[0.0, 0.0, 0.0, 1.0, 1.0, 1.0]], device=device)
    near_planes = torch.tensor([1.2], device=device)
    far_planes = torch.tensor([1.5], device=device)
    intervals, samples, _ = traverse_grids(rays_o=rays_o, rays_d=rays_d, binaries=binaries, aabbs=aabbs,
                                           step_size=0.05, near_planes=near_planes, far_planes=far_planes)
    assert intervals.vals.numel() > 0
    assert (intervals.vals >= (near_planes - 0.05 / 2)).all()
    assert (intervals.vals <= (far_planes + 0.05 / 2)).all()


def test_sampling_with_min_max_distances(apnerf):
    from apnerf.nerfacc import OccGridEstimator

    torch.manual_seed(42)
    n_rays, levels, resolution, step = 64, 4, 32, 0.01
    rays_o = torch.rand((n_rays, 3), device=device) * 2 - 1.0
    rays_d = torch.rand((n_rays, 3), device=device)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], device=device)
    binaries = torch.rand((levels, resolution, resolution, resolution), device=device) > 0.5
    t_min = torch.rand((n_rays,), device=device)
    t_max = t_min + torch.rand((n_rays,), device=device)
    est = OccGridEstimator(roi_aabb=aabb, resolution=resolution, levels=levels).to(device)
    est.binaries = binaries
    ray_indices, t_starts, t_ends = est.sampling(rays_o=rays_o, rays_d=rays_d, near_plane=0.15, far_plane=0.85,
                                                 t_min=t_min, t_max=t_max, render_step_size=step)
    assert t_starts.numel() > 0
    assert (t_starts >= (t_min[ray_indices] - step / 2)).all()
    assert (t_ends <= (t_max[ray_indices] + step / 2)).all()


def test_render_visibility_and_alpha_weights(apnerf):
    from apnerf.nerfacc.volrend import render_visibility_from_alpha, render_weight_from_alpha

    ray_indices = torch.tensor([0, 2, 2, 2, 2], dtype=torch.int64, device=device)
    alphas = torch.tensor([0.4, 0.3, 0.8, 0.8, 0.5], dtype=torch.float32, device=device)
    vis = render_visibility_from_alpha(alphas, ray_indices=ray_indices, n_rays=3, early_stop_eps=0.03, alpha_thre=0.0)
    assert vis.tolist() == [True, True, True, True, False]
    vis = render_visibility_from_alpha(alphas, ray_indices=ray_indices, n_rays=3, early_stop_eps=0.05, alpha_thre=0.35)
    assert vis.tolist() == [True, False, True, True, False]
    weights, _ = render_weight_from_alpha(alphas, ray_indices=ray_indices, n_rays=3)
    tgt = torch.tensor([1.0 * 0.4, 1.0 * 0.3, 0.7 * 0.8, 0.14 * 0.8, 0.028 * 0.5], device=device)
    assert torch.allclose(weights, tgt, atol=1e-6)


def test_render_weight_from_density(apnerf):
    from apnerf.nerfacc.volrend import render_weight_from_alpha, render_weight_from_density

    ray_indices = torch.tensor([0, 2, 2, 2, 2], dtype=torch.int64, device=device)
    sigmas = torch.rand((5,), device=device)
    t_starts = torch.rand_like(sigmas)
    t_ends = torch.rand_like(sigmas) + 1.0
    alphas = 1.0 - torch.exp(-sigmas * (t_ends - t_starts))
    weights, _, _ = render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=3)
    weights_tgt, _ = render_weight_from_alpha(alphas, ray_indices=ray_indices, n_rays=3)
    assert torch.allclose(weights, weights_tgt, atol=1e-6)


def test_accumulate_along_rays(apnerf):
    from apnerf.nerfacc import accumulate_along_rays

    ray_indices = torch.tensor([0, 2, 2, 2, 2], dtype=torch.int64, device=device)
    weights = torch.tensor([0.4, 0.3, 0.8, 0.8, 0.5], dtype=torch.float32, device=device)
    for D in (2, 29):
        values = torch.rand((5, D), device=device)
        ray_values = accumulate_along_rays(weights, values=values, ray_indices=ray_indices, n_rays=3)
        assert ray_values.shape == (3, D)
        assert torch.allclose(ray_values[0, :], weights[0, None] * values[0, :])
        assert (ray_values[1, :] == 0).all()
        assert torch.allclose(ray_values[2, :], (weights[1:, None] * values[1:]).sum(dim=0))


def test_grads(apnerf):
    """tests/test_rendering.py:110-193 golden weights and sigma-gradients (atol 1e-4)."""
    from apnerf.nerfacc.volrend import render_transmittance_from_density, render_weight_from_density

    ray_indices = torch.tensor([0, 2, 2, 2, 2], dtype=torch.int64, device=device)
    packed_info = torch.tensor([[0, 1], [1, 0], [1, 4]], dtype=torch.long, device=device)
    sigmas = torch.tensor([0.4, 0.8, 0.1, 0.8, 0.1], device=device, requires_grad=True)
    t_starts = torch.rand(5, device=device)
    t_ends = t_starts + 1.0
    weights_ref = torch.tensor([0.3297, 0.5507, 0.0428, 0.2239, 0.0174], device=device)
    grad_ref = torch.tensor([0.6703, 0.1653, 0.1653, 0.1653, 0.1653], device=device)
    for kw in (dict(ray_indices=ray_indices, n_rays=3), dict(packed_info=packed_info, n_rays=3)):
        trans, _ = render_transmittance_from_density(t_starts, t_ends, sigmas, **kw)
        weights = trans * (1.0 - torch.exp(-sigmas * (t_ends - t_starts)))
        weights.sum().backward()
        g = sigmas.grad.clone()
        sigmas.grad.zero_()
        assert torch.allclose(weights_ref, weights, atol=1e-4)
        assert torch.allclose(grad_ref, g, atol=1e-4), g
        weights, _, _ = render_weight_from_density(t_starts, t_ends, sigmas, **kw)
        weights.sum().backward()
        g = sigmas.grad.clone()
        sigmas.grad.zero_()
        assert torch.allclose(weights_ref, weights, atol=1e-4)
        assert torch.allclose(grad_ref, g, atol=1e-4), g


def test_rendering_and_pack_info(apnerf):
    from apnerf.nerfacc import pack_info, rendering

    ray_indices = torch.tensor([0, 2, 2, 2, 2], dtype=torch.int64, device=device)
    assert pack_info(ray_indices, n_rays=3).tolist() == [[0, 1], [1, 0], [1, 4]]
    sigmas = torch.rand((5,), device=device)
    t_starts = torch.rand_like(sigmas)
    t_ends = torch.rand_like(sigmas) + 1.0
    colors, opac, depth, extras = rendering(
        t_starts, t_ends, ray_indices=ray_indices, n_rays=3,
        rgb_sigma_fn=lambda ts, te, ri: (torch.stack([ts] * 3, dim=-1), ts))
    assert colors.shape == (3, 3) and opac.shape == (3, 1) and depth.shape == (3, 1)


def test_packed_scans_match_cumsum(apnerf):
    from apnerf.nerfacc import exclusive_sum, inclusive_sum

    torch.manual_seed(42)
    for fn, atol in ((inclusive_sum, 1e-4), (exclusive_sum, 3e-4)):
        data = torch.rand((5, 1000), device=device, requires_grad=True)
        out1 = fn(data).flatten()
        out1.sum().backward()
        grad1 = data.grad.clone()
        data.grad.zero_()
        chunk_starts = torch.arange(0, data.numel(), data.shape[1], device=device, dtype=torch.long)
        chunk_cnts = torch.full((data.shape[0],), data.shape[1], dtype=torch.long, device=device)
        out2 = fn(data.flatten(), packed_info=torch.stack([chunk_starts, chunk_cnts], dim=-1))
        out2.sum().backward()
        assert torch.allclose(out1, out2, atol=atol)
        assert torch.allclose(grad1, data.grad, atol=1e-3, rtol=1e-5)


def test_volrend_integration_config1(apnerf, oracle):
    """BASELINE.json configs[0]: 4096 rays x <= 1024 packed samples, rgb / depth / 29-class
    semantics, fp32.  Tolerance: <= 1e-5 relative (fp32 compositing)."""
    from apnerf.nerfacc import accumulate_along_rays, render_weight_from_density

    rng = np.random.default_rng(0)
    R = 4096
    cnts = rng.integers(0, 1025, R)
    ray_indices = np.repeat(np.arange(R), cnts)
    N = ray_indices.size
    dt = (1e-3 * (1 + rng.random(N))).astype(np.float32)
    starts = np.cumsum(cnts) - cnts
    csum = np.cumsum(dt, dtype=np.float64)
    t_starts = (0.1 + csum - dt - np.repeat(np.concatenate([[0], csum])[starts], cnts)).astype(np.float32)
    t_ends = t_starts + dt
    sigmas = (np.exp(rng.standard_normal(N)) * 20).astype(np.float32)
    rgbs = rng.random((N, 3)).astype(np.float32)
    sems = (2 * rng.standard_normal((N, 29))).astype(np.float32)
    ow, ot, oa = oracle.render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=R)
    # float64 ground truth of the same formulas (volrend.py:259-267)
    sdt64 = sigmas.astype(np.float64) * (t_ends.astype(np.float64) - t_starts.astype(np.float64))
    cs = np.cumsum(sdt64)
    ex64 = cs - sdt64 - np.repeat(np.concatenate([[0.0], cs])[starts], cnts)
    t64 = np.exp(-ex64)
    a64 = 1.0 - np.exp(-sdt64)
    w64 = t64 * a64
    tt = lambda a: torch.from_numpy(a).to(device)
    ri = tt(ray_indices)
    w, t, a = render_weight_from_density(tt(t_starts), tt(t_ends), tt(sigmas), ray_indices=ri, n_rays=R)

    def rel_err(x, y, floor=1e-3):
        x = np.asarray(x.cpu().numpy() if torch.is_tensor(x) else x, np.float64)
        return (np.abs(x - y) / np.maximum(np.abs(y), floor * np.abs(y).max())).max()

    def abs_err(x, y):
        return np.abs(np.asarray(x.cpu().numpy(), np.float64) - y).max()

    # per-sample quantities live in [0, 1].  alpha = 1 - exp(-sigma*dt) cancels in fp32 (in the reference
    # too, volrend.py:260), so a small alpha carries up to ~1 ulp(1.0) = 6e-8 ABSOLUTE error whichever exp()
    # is used: the per-sample bar is absolute, the relative 1e-5 bar applies to the rendered outputs below.
    assert abs_err(w, w64) <= 2e-7 and abs_err(t, t64) <= 1e-6 and abs_err(a, a64) <= 2e-7, \
        (abs_err(w, w64), abs_err(t, t64), abs_err(a, a64))
    assert abs_err(w, ow.astype(np.float64)) <= 2e-7 and rel_err(t, ot.astype(np.float64)) <= 3e-4
    # rendered outputs (what the north star bounds: fp32 compositing <= 1e-5 relative), CUDA weights and
    # CUDA accumulation against the float64 ground truth
    tmid32 = (tt(t_starts) + tt(t_ends))[:, None] / 2.0
    tmid64 = tmid32.cpu().numpy().astype(np.float64)
    for vals, v64, name in ((tt(rgbs), rgbs.astype(np.float64), "rgb"), (None, None, "opacity"),
                            (tmid32, tmid64, "depth"), (tt(sems), sems.astype(np.float64), "sem")):
        got = accumulate_along_rays(w, vals, ri, R).cpu().numpy().astype(np.float64)
        src = w64[:, None] * (1.0 if v64 is None else v64)
        exp = np.zeros((R, src.shape[1]))
        np.add.at(exp, ray_indices, src)
        mag = np.zeros_like(exp)  # sum of |terms|: the scale a summation error is relative to
        np.add.at(mag, ray_indices, np.abs(src))
        err = np.abs(got - exp) / np.maximum(mag, 1e-2 * mag.max())
        assert err.max() <= 1e-5, (name, err.max())
