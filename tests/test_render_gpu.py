"""GPU: the device-driven test-mode renderer and the scorer.

* marching schedule + ray-march samples: bit-exact against the CPU oracle's restatement of
  perception/models/utils.py:782-1032 on a low-density field (no ray saturates, so the
  schedule is purely geometric and must agree iteration by iteration);
* rendered outputs on the "trained-like" field: fused == op-by-op == oracle within the
  north-star tolerances (written at each assert);
* predictive information: fp32 per-pixel entropies with float64 sums against the float64 oracle,
  1e-5 absolute on each term (the north star allows 1e-3 on entropy).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OPTS = dict(near_plane=0.1, render_step_size=1e-3, cone_angle=0.004, alpha_thre=0.01)


def _scene(apnerf, density_gain, C=29, res=128):
    from apnerf import synthetic

    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=res, levels=1)
    est.binaries = synthetic.make_occupancy(res, seed=1)
    field = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=C)
    synthetic.init_trained_like(field, seed=2, density_gain=density_gain)
    return field.to(DEV).eval(), est.to(DEV).eval()


def _rays(apnerf, w, h, pose_seed=3):
    from apnerf import synthetic
    from oracle import oracle as O

    pose = synthetic.pose_to_matrix(synthetic.make_poses_corridor(1, seed=pose_seed)[0]).astype(np.float32)
    o, d = O.generate_image_rays(pose, w, h, focal=w / 2)
    return o, d, pose


def _oracle_field(oracle, field):
    fp = oracle.FieldParams(field.mlp_base.params.detach().cpu().numpy(), field.mlp_head.params.detach().cpu().numpy(),
                            field.mlp_sem.params.detach().cpu().numpy(), num_semantic_classes=29)
    aabb = field.aabb.cpu().numpy()
    return lambda p, d: oracle.field_forward(p, d, aabb, fp)


def test_generate_rays_matches_reference_formula(apnerf, oracle):
    from apnerf import synthetic
    from apnerf._lib import call

    poses = synthetic.make_poses_corridor(3, seed=9)
    c2w = torch.from_numpy(apnerf.scoring.poses_to_c2w(poses)).to(DEV)
    w, h = 64, 48
    ro = torch.empty((3 * w * h, 3), device=DEV)
    rd = torch.empty((3 * w * h, 3), device=DEV)
    call("apnerf_generate_rays", 3, c2w, w, h, float(w / 2), w * h, None, ro, rd)
    for v in range(3):
        o, d = oracle.generate_image_rays(synthetic.pose_to_matrix(poses[v]).astype(np.float32), w, h, w / 2)
        assert np.array_equal(ro[v * w * h:(v + 1) * w * h].cpu().numpy(), o)
        assert np.abs(rd[v * w * h:(v + 1) * w * h].cpu().numpy() - d).max() <= 2e-7
    # rounded-linspace subsample (habitat_to_data.py:462-467)
    keep = oracle.subsample_indices(w * h, 100).astype(np.int32)
    ro2 = torch.empty((100, 3), device=DEV)
    rd2 = torch.empty((100, 3), device=DEV)
    call("apnerf_generate_rays", 1, c2w[:1].contiguous(), w, h, float(w / 2), 100, torch.from_numpy(keep).to(DEV), ro2, rd2)
    assert torch.equal(rd2, rd[:w * h][torch.from_numpy(keep).long().to(DEV)])


@pytest.mark.parametrize("fuse", [False, True])
def test_schedule_and_samples_bit_exact(apnerf, oracle, fuse):
    """density ~ e^-1 everywhere -> alpha < alpha_thre, opacity stays 0: every implementation
    must take exactly the same marching decisions."""
    field, est = _scene(apnerf, density_gain=0.0)
    w, h = 40, 30
    o, d, _ = _rays(apnerf, w, h)
    trace = []
    oracle.render_probablistic_image_with_occgrid_test(
        256, _oracle_field(oracle, field), est.binaries.cpu().numpy(), est.aabbs.cpu().numpy(), o, d, 29,
        trace=trace, **OPTS)
    got = []

    def hook(it, r):
        n = int(r.counters[2].item())
        ray = r.s_ray[:n].cpu().numpy().astype(np.int64)
        ts, te = r.s_ts[:n].cpu().numpy(), r.s_te[:n].cpu().numpy()
        if fuse:  # tile-aware layout: padding rows carry ray = -1 and no ray straddles a 128-row tile
            real = ray >= 0
            cnt = r.s_cnt[:n].cpu().numpy()
            heads = np.nonzero((cnt > 0) & (cnt < 128))[0]  # k at a ray's first row; 0x80 | j on its other rows
            assert ((heads % 128) + cnt[heads] <= 128).all(), "a ray's samples must not straddle a tile"
            ray, ts, te = ray[real], ts[real], te[real]
        order = np.lexsort((ts, ray))
        got.append(dict(n_live=int(r.counters[0].item()), n_samples=int(r.n_samp[0].item()), ray=ray[order],
                        ts=ts[order], te=te[order]))

    r = apnerf.FusedRenderer(DEV, 29)
    r.render(field, est, torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), w * h, max_samples=256,
             poll_every=0, debug_hook=hook, fuse_compositor=fuse, **OPTS)
    assert len(trace) >= 10 and sum(len(t["ray_indices"]) for t in trace) > 10000
    for it, t in enumerate(trace):
        g = got[it]
        assert g["n_live"] == t["n_alive"] and g["n_samples"] == t["n_samples"], (it, g["n_live"], t["n_alive"])
        assert np.array_equal(g["ray"], t["ray_indices"]), f"iteration {it}: ray indices"
        assert np.array_equal(g["ts"].view(np.int32), t["t_starts"].view(np.int32)), f"iteration {it}: t_starts"
        assert np.array_equal(g["te"].view(np.int32), t["t_ends"].view(np.int32)), f"iteration {it}: t_ends"
    for g in got[len(trace):]:
        assert g["n_live"] == 0 or g["n_samples"] == 0


def test_fused_render_matches_unfused_and_oracle(apnerf, oracle):
    field, est = _scene(apnerf, density_gain=6.0)
    w, h = 48, 36
    o, d, _ = _rays(apnerf, w, h)
    rays = apnerf.Rays(origins=torch.from_numpy(o).to(DEV), viewdirs=torch.from_numpy(d).to(DEV))
    bk = torch.zeros(3, device=DEV)
    fused = apnerf.render_probablistic_image_with_occgrid_test(1024, field, est, rays, render_bkgd=bk, **OPTS)
    # the variant with the compositor fused into the field kernel's epilogue
    r4 = apnerf.FusedRenderer(DEV, 29)
    st4 = r4.render(field, est, rays.origins, rays.viewdirs, w * h, max_samples=1024, fuse_compositor=True, **OPTS)
    r4.check_overflow()
    o4 = r4.finalize(st4, bk)
    for key, a in zip(("rgb", "rgb_var", "opacity", "depth", "depth_var", "sem"), fused[:6]):
        assert torch.allclose(a, o4[key], rtol=0, atol=1e-6 * max(1.0, float(o4[key].abs().max()))), key
    unfused = apnerf.render.render_probablistic_image_with_occgrid_test_unfused(1024, field, est, rays, render_bkgd=bk,
                                                                               **OPTS)
    orc = oracle.render_probablistic_image_with_occgrid_test(
        1024, _oracle_field(oracle, field), est.binaries.cpu().numpy(), est.aabbs.cpu().numpy(), o, d, 29, **OPTS)
    names = ["rgb", "rgb_var", "opacity", "depth", "depth_var", "sem"]
    opac = orc[2]
    assert 0.05 < (opac > 0.5).mean() < 1.0, "scene should have both saturated and open rays"
    assert abs(fused[6] - orc[6]) <= 0.02 * orc[6] and abs(unfused[6] - orc[6]) <= 0.02 * orc[6]
    from bounds import assert_bounded, scale_of

    vs_unfused, vs_oracle = [], []
    for name, a, b, c in zip(names, fused[:6], unfused[:6], orc[:6]):
        a, b = a.cpu().numpy(), b.cpu().numpy()
        vs_unfused.append((name, a, b, 1e-5 * scale_of(c)))
        vs_oracle.append((name, a, c, 1e-3 * scale_of(c)))
    # fused vs op-by-op on the same GPU: the field kernel's two entry points (sample rows / explicit points) evaluate the
    # same network; the fp32 accumulation order of the compositing differs: 1e-5 of the range on every pixel but a
    # counted handful whose samples sit on a threshold
    assert_bounded(vs_unfused, w * h, "fused vs op-by-op")
    # vs the CPU oracle: fp16 MLP tolerance 1e-3 of the range (north star) on every pixel except counted flip pixels
    assert_bounded(vs_oracle, w * h, "fused vs oracle")


def test_plain_renderer_and_no_semantics(apnerf):
    field, est = _scene(apnerf, density_gain=6.0, C=0)
    w, h = 32, 24
    o, d, _ = _rays(apnerf, w, h)
    rays = apnerf.Rays(origins=torch.from_numpy(o).to(DEV).view(h, w, 3), viewdirs=torch.from_numpy(d).to(DEV).view(h, w, 3))
    rgb, opacity, depth, total = apnerf.render_image_with_occgrid_test(1024, field, est, rays,
                                                                        render_bkgd=torch.ones(3, device=DEV), **OPTS)
    assert rgb.shape == (h, w, 3) and opacity.shape == (h, w, 1) and depth.shape == (h, w, 1) and total > 0
    p = apnerf.render_probablistic_image_with_occgrid_test(1024, field, est, rays, render_bkgd=torch.ones(3, device=DEV),
                                                           **OPTS)
    assert len(p) == 6 and torch.allclose(p[0], rgb, atol=1e-6) and torch.allclose(p[2], opacity, atol=1e-6)


def test_predictive_information_matches_oracle(apnerf, oracle):
    from apnerf._lib import call

    g = torch.Generator().manual_seed(0)
    V, R, C, E, T = 6, 500, 29, 2, 3
    NR = V * R
    states = []
    for m in range(E):
        st = torch.zeros((9 + C, NR))
        st[5:8] = torch.rand((3, NR), generator=g) * 0.05
        st[8] = torch.rand(NR, generator=g) * 0.5
        st[3] = torch.rand(NR, generator=g)
        st[9:] = torch.randn((C, NR), generator=g) * 3
        states.append(st.to(DEV))
    view_traj = torch.tensor([0, 0, 1, 2, 2, 2], dtype=torch.int32, device=DEV)
    sums = torch.zeros((T, 4), dtype=torch.float64, device=DEV)
    call("apnerf_score_views", E, states[0], states[1], None, None, NR, R, C, view_traj, T, sums)
    counts = np.array([2, 1, 3]) * R
    terms = apnerf.PredictiveInformationScorer.finish(sums.cpu().numpy(), counts)
    vt = view_traj.cpu().numpy()
    for t in range(T):
        sel = np.repeat(vt == t, R)
        st = [s.cpu().numpy()[:, sel] for s in states]
        ref = oracle.predictive_information(
            np.stack([s[5:8].T for s in st]), np.stack([s[8] for s in st]), np.stack([s[3] for s in st]),
            np.stack([s[9:].T for s in st]))
        # per-pixel entropies in fp32 (<= 2 ulp logf / expf), sums in float64: 1e-5 absolute on the means
        assert np.abs(terms[t] - ref).max() <= 1e-5, (t, terms[t], ref)


def test_score_trajectories_end_to_end(apnerf, oracle):
    """poses -> scores through the public scorer, against the oracle chain (CPU render of every
    view by both members + float64 scoring).  Tolerance 1e-3 absolute on each entropy term."""
    from apnerf import synthetic

    f0, e0 = _scene(apnerf, density_gain=6.0)
    f1 = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=29)
    synthetic.init_trained_like(f1, seed=12, density_gain=6.0)
    f1 = f1.to(DEV).eval()
    w, h = 24, 18
    scorer = apnerf.PredictiveInformationScorer([f0, f1], [e0, e0], w, h, w / 2, views_per_batch=3, **OPTS)
    poses = synthetic.make_poses_corridor(5, seed=21)
    view_traj = np.array([0, 0, 0, 1, 1], dtype=np.int32)
    terms = scorer.score_views(poses, view_traj, 2)
    outs = [[], []]
    for m, f in enumerate((f0, f1)):
        fn = _oracle_field(oracle, f)
        for v in range(5):
            o, d = oracle.generate_image_rays(synthetic.pose_to_matrix(poses[v]).astype(np.float32), w, h, w / 2)
            outs[m].append(oracle.render_probablistic_image_with_occgrid_test(
                1024, fn, e0.binaries.cpu().numpy(), e0.aabbs.cpu().numpy(), o, d, 29, **OPTS))
    for t in range(2):
        vs = np.nonzero(view_traj == t)[0]
        stack = lambda k: np.stack([np.stack([outs[m][v][k] for v in vs]) for m in range(2)])
        ref = oracle.predictive_information(stack(1), stack(4)[..., 0], stack(2)[..., 0], stack(5))
        assert np.abs(terms[t] - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max()), (terms[t], ref)
