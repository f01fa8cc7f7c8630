"""GPU: edge cases the reference guards explicitly -- empty and ragged inputs, rays that miss the
grid, fully empty / fully occupied grids, zero live rays, masks that disable everything."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_empty_inputs_return_empty(apnerf):
    from apnerf import nerfacc

    e = torch.empty(0, device=DEV)
    ei = torch.empty(0, dtype=torch.int64, device=DEV)
    pi = torch.zeros((3, 2), dtype=torch.int64, device=DEV)
    w, t, a = nerfacc.render_weight_from_density(e, e, e, packed_info=pi)      # scan.cu:32-34
    assert w.numel() == t.numel() == a.numel() == 0
    assert nerfacc.exclusive_sum(e, pi).numel() == 0
    out = nerfacc.accumulate_along_rays(e, torch.empty((0, 29), device=DEV), ei, 5)
    assert out.shape == (5, 29) and (out == 0).all()
    assert nerfacc.pack_info(ei, 4).tolist() == [[0, 0]] * 4
    o = torch.empty((0, 3), device=DEV)
    tm, tx, h = nerfacc.ray_aabb_intersect(o, o, torch.tensor([[0.0, 0, 0, 1, 1, 1]], device=DEV))
    assert tm.shape == (0, 1)


def test_rays_missing_the_grid_and_empty_grid(apnerf):
    from apnerf import nerfacc

    aabbs = torch.tensor([[0.0, 0, 0, 1, 1, 1]], device=DEV)
    rays_o = torch.tensor([[5.0, 5, 5], [-1.0, 0.5, 0.5], [0.5, 0.5, 0.5]], device=DEV)
    rays_d = torch.tensor([[1.0, 0, 0], [1.0, 0, 0], [0.0, 0, 1.0]], device=DEV)
    full = torch.ones((1, 8, 8, 8), dtype=torch.bool, device=DEV)
    iv, sm, term = nerfacc.traverse_grids(rays_o, rays_d, full, aabbs, step_size=0.05)
    cnt = sm.packed_info[:, 1].tolist()
    assert cnt[0] == 0 and cnt[1] > 0 and cnt[2] > 0          # miss, through, starting inside
    assert (sm.ray_indices >= 1).all()
    assert iv.is_left.sum() == iv.is_right.sum() == sm.vals.numel()
    empty = torch.zeros_like(full)
    iv, sm, _ = nerfacc.traverse_grids(rays_o, rays_d, empty, aabbs, step_size=0.05)
    assert sm.vals.numel() == 0 and iv.vals.numel() == 0 and sm.packed_info[:, 1].sum() == 0
    # everything masked in test mode: nothing is written, counts stay 0
    mask = torch.zeros(3, dtype=torch.bool, device=DEV)
    iv, sm, _ = nerfacc.traverse_grids(rays_o, rays_d, full, aabbs, step_size=0.05, traverse_steps_limit=8,
                                       over_allocate=True, rays_mask=mask)
    assert sm.packed_info[:, 1].sum() == 0 and sm.is_valid.sum() == 0


def test_ragged_segments_including_empty_rays(apnerf, oracle):
    from apnerf import nerfacc

    rng = np.random.default_rng(5)
    cnts = np.array([0, 1, 0, 33, 64, 0, 1000, 31, 0], dtype=np.int64)
    starts = np.cumsum(cnts) - cnts
    n = int(cnts.sum())
    x = rng.random(n).astype(np.float32)
    pi = torch.from_numpy(np.stack([starts, cnts], -1)).to(DEV)
    got = nerfacc.exclusive_sum(torch.from_numpy(x).to(DEV), pi).cpu().numpy()
    ref = oracle.exclusive_sum(x, np.stack([starts, cnts], -1))
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)
    ri = np.repeat(np.arange(len(cnts)), cnts)
    assert nerfacc.pack_info(torch.from_numpy(ri).to(DEV), len(cnts)).cpu().numpy().tolist() == np.stack([starts, cnts], -1).tolist()
    # accumulate with UNSORTED ray indices (index_add_ semantics do not require packing)
    perm = rng.permutation(n)
    vals = rng.random((n, 3)).astype(np.float32)
    got = nerfacc.accumulate_along_rays(torch.from_numpy(x[perm]).to(DEV), torch.from_numpy(vals[perm]).to(DEV),
                                        torch.from_numpy(ri[perm]).to(DEV), len(cnts)).cpu().numpy()
    ref = oracle.accumulate_along_rays(x, vals, ri, len(cnts))
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-4)


def test_render_view_that_sees_nothing(apnerf):
    """All rays miss the aabb: zero samples, opacity 0, rgb = background, depth 0, finite outputs."""
    from apnerf import synthetic

    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
    est.binaries = synthetic.make_occupancy(128, seed=1)
    est = est.to(DEV).eval()
    f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=29)
    f = synthetic.init_trained_like(f).to(DEV).eval()
    n = 500
    rays = apnerf.Rays(origins=torch.tensor([100.0, 100.0, 100.0], device=DEV).expand(n, 3).contiguous(),
                       viewdirs=torch.tensor([1.0, 0.0, 0.0], device=DEV).expand(n, 3).contiguous())
    bk = torch.tensor([0.25, 0.5, 0.75], device=DEV)
    rgb, rgb_var, acc, depth, depth_var, sem, total = apnerf.render_probablistic_image_with_occgrid_test(
        1024, f, est, rays, near_plane=0.1, render_step_size=1e-3, render_bkgd=bk, cone_angle=0.004, alpha_thre=0.01)
    assert total == 0 and (acc == 0).all() and (depth == 0).all() and (sem == 0).all() and (rgb_var == 0).all()
    assert torch.allclose(rgb, bk.expand(n, 3))


def test_field_rejects_unsupported_configs_loudly(apnerf):
    with pytest.raises(NotImplementedError):
        apnerf.NGPRadianceField([-1, -1, -1, 1, 1, 1], layers=2, neurons=64)
    with pytest.raises(NotImplementedError):
        apnerf.NGPRadianceField([-1, -1, -1, 1, 1, 1], layers=2, unbounded=True)
    f = apnerf.NGPRadianceField([-1, -1, -1, 1, 1, 1], layers=2, num_semantic_classes=3)
    with pytest.raises(RuntimeError):
        f.query_density(torch.zeros(4, 3))  # CPU tensors: no fallback


def test_closed_form_skip_is_bit_exact_when_forced():
    """The exact closed-form empty-space skip (csrc/march.cuh: skip_to) is only taken for skips of more than
    APNERF_SKIP_MIN = 256 steps by default, which the ordinary fixtures rarely reach.  Re-run the bit-exact
    comparisons against the reference's own kernels and the oracle in a child process that forces it for every
    skip longer than two steps (the knob is read once per process)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, APNERF_SKIP_MIN="2")
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        os.path.join(root, "tests", "test_gpu_reference.py"),
                        os.path.join(root, "tests", "test_render_gpu.py") + "::test_schedule_and_samples_bit_exact"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
