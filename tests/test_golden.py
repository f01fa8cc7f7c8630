"""Golden vectors from the reference's own Python (tests/golden/make_golden.py, run once where
/root/reference exists) against (a) the CPU oracle -- this pins the oracle's restatement of
utils.py:783-1032, habitat_to_data.py:274-301/413-548 and pipeline.py:666-798 -- and (b) the CUDA path.

What the fixture cannot pin is the tiny-cuda-nn field (absent from /root/reference): both the generator
and the tests evaluate it with the oracle's restatement, see the header of the generator.
"""
import os

import numpy as np
import pytest
import torch

from bounds import assert_bounded, scale_of

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_python_path.npz")
NAMES = ("rgb", "rgb_var", "opacity", "depth", "depth_var", "sem")


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLDEN)
    cfg = {k: g[f"cfg_{k}"].tolist() for k in g["cfg_keys"]}
    return g, cfg


def _members(apnerf, cfg, device="cpu"):
    from apnerf import synthetic

    out = []
    for occ_seed, field_seed in zip(cfg["occ_seeds"], cfg["field_seeds"]):
        est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=cfg["grid_res"], levels=1)
        est.binaries = synthetic.make_occupancy(cfg["grid_res"], seed=occ_seed)
        f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=cfg["n_classes"])
        synthetic.init_trained_like(f, seed=field_seed, density_gain=cfg["density_gain"])
        out.append((f.to(device).eval(), est.to(device).eval()))
    return out


def _oracle_field(O, field, C):
    fp = O.FieldParams(field.mlp_base.params.detach().cpu().numpy(), field.mlp_head.params.detach().cpu().numpy(),
                       field.mlp_sem.params.detach().cpu().numpy(), num_semantic_classes=C)
    aabb = field.aabb.cpu().numpy()
    return lambda p, d: O.field_forward(p, d, aabb, fp)


VIEW_OPTS = {
    "a": dict(max_samples=1024, cone_angle=0.004, alpha_thre=0.01, bkgd=(0.0, 0.0, 0.0), step_mul=1),
    "b": dict(max_samples=96, cone_angle=0.0, alpha_thre=0.0, bkgd=(0.2, 0.5, 0.9), step_mul=8),
}


# ------------------------------------------------------------------------------------------------
# CPU: oracle == reference Python
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_renderer_matches_reference_python(apnerf, oracle, gold, tag):
    """Same native stand-ins and the same field on both sides, so the only differences allowed are numpy vs
    torch fp32 elementwise rounding (exp, pow): 1e-6 relative to the largest value of each output, and the
    sample count must be identical."""
    g, cfg = gold
    (field, est), _ = _members(apnerf, cfg)
    o = VIEW_OPTS[tag]
    got = oracle.render_probablistic_image_with_occgrid_test(
        o["max_samples"], _oracle_field(oracle, field, cfg["n_classes"]), est.binaries.numpy(), est.aabbs.numpy(),
        g["view_rays_o"], g["view_rays_d"], cfg["n_classes"], near_plane=cfg["near_plane"],
        render_step_size=cfg["render_step_size"] * o["step_mul"], render_bkgd=np.asarray(o["bkgd"], np.float32),
        cone_angle=o["cone_angle"], alpha_thre=o["alpha_thre"])
    assert int(got[6]) == int(g[f"view_{tag}_total_samples"])
    for name, a in zip(NAMES, got[:6]):
        ref = g[f"view_{tag}_{name}"]
        assert a.shape == ref.shape, name
        assert np.abs(a - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), (name, np.abs(a - ref).max())


def test_oracle_rays_and_view_selection_match_reference_python(apnerf, oracle, gold):
    """generate_image_rays + the rounded-linspace subsample (habitat_to_data.py:274-301, 462-467) and the
    trajectory's view indices (pipeline.py:688-690)."""
    from apnerf import synthetic

    g, cfg = gold
    traj = g["traj_poses"]
    pose = synthetic.pose_to_matrix(traj[3]).astype(np.float32)
    o, d = oracle.generate_image_rays(pose, cfg["img_w"], cfg["img_h"], cfg["img_w"] / 2.0)
    idx = oracle.subsample_indices(o.shape[0], 24 * 32)
    assert np.array_equal(o[idx], g["view_rays_o"])
    assert np.abs(d[idx] - g["view_rays_d"]).max() <= 2e-7  # torch.linalg.norm vs explicit sqrt of the sum
    unc = apnerf.scoring.uncertainty_view_indices(len(traj))
    a = np.linspace(0, len(traj) - 20, 20)
    b = np.linspace(len(traj) - 20, len(traj) - 1, 20)
    assert np.array_equal(unc, np.hstack((a, b)).astype(int))


def test_oracle_predictive_information_matches_reference_python(oracle, gold):
    """The four terms of pipeline.py:727-790 from the reference's own renders of the trajectory."""
    g, cfg = gold
    E = cfg["n_ensembles"]
    stack = lambda k: np.stack([g[f"traj_m{m}_{k}"] for m in range(E)]).astype(np.float64)
    terms = oracle.predictive_information(stack("rgb_var"), stack("depth_var"), stack("acc"), stack("sem"))
    assert np.abs(terms - g["traj_terms"]).max() <= 1e-9, (terms, g["traj_terms"])
    assert abs(terms.sum() - float(g["traj_score"])) <= 1e-9


def test_oracle_trajectory_views_match_reference_python(apnerf, oracle, gold):
    """A few of the trajectory's views through the oracle chain pose -> rays -> subsample -> render, against the
    images Dataset.render_probablistic_image_from_pose produced (bounded: 3 views of member 1)."""
    from apnerf import synthetic

    g, cfg = gold
    members = _members(apnerf, cfg)
    field, est = members[1]
    traj = g["traj_poses"]
    unc = apnerf.scoring.uncertainty_view_indices(len(traj))
    w, h = cfg["img_w"], cfg["img_h"]
    sw, sh = int(w * cfg["scale"]), int(h * cfg["scale"])
    fn = _oracle_field(oracle, field, cfg["n_classes"])
    for vi in (0, 19, 39):
        pose = synthetic.pose_to_matrix(traj[unc[vi]]).astype(np.float32)
        o, d = oracle.generate_image_rays(pose, w, h, w / 2.0)
        idx = oracle.subsample_indices(o.shape[0], sw * sh)
        r = oracle.render_probablistic_image_with_occgrid_test(
            1024, fn, est.binaries.numpy(), est.aabbs.numpy(), o[idx], d[idx], cfg["n_classes"],
            near_plane=cfg["near_plane"], render_step_size=cfg["render_step_size"], cone_angle=cfg["cone_angle"],
            alpha_thre=cfg["alpha_thre"])
        for name, k in (("rgb", 0), ("rgb_var", 1), ("acc", 2), ("depth", 3), ("depth_var", 4), ("sem", 5)):
            ref = g[f"traj_m1_{name}"][vi]
            a = r[k].reshape(ref.shape)
            # the ray directions differ by <= 2e-7 (norm rounding), which moves sample positions by ~1e-6 m
            assert np.quantile(np.abs(a - ref), 0.99) <= 2e-4 * max(1.0, np.abs(ref).max()), (vi, name)


# ------------------------------------------------------------------------------------------------
# GPU: CUDA path == reference Python (through the reference-facing calls)
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_cuda_renderer_matches_reference_python(apnerf, gold, tag):
    """render_probablistic_image_with_occgrid_test (drop-in signature) on the golden rays.  Tolerances are the
    north star's: every pixel of every output within 1e-3 of the output's range (fp16 MLP tolerance), except an
    explicitly counted handful of threshold-flip pixels (tests/bounds.py)."""
    g, cfg = gold
    dev = "cuda:0"
    (field, est), _ = _members(apnerf, cfg, dev)
    o = VIEW_OPTS[tag]
    rays = apnerf.Rays(origins=torch.from_numpy(g["view_rays_o"]).to(dev), viewdirs=torch.from_numpy(g["view_rays_d"]).to(dev))
    got = apnerf.render_probablistic_image_with_occgrid_test(
        o["max_samples"], field, est, rays, near_plane=cfg["near_plane"],
        render_step_size=cfg["render_step_size"] * o["step_mul"], render_bkgd=torch.tensor(o["bkgd"], device=dev),
        cone_angle=o["cone_angle"], alpha_thre=o["alpha_thre"])
    ref_total = int(g[f"view_{tag}_total_samples"])
    assert abs(int(got[6]) - ref_total) <= 0.02 * ref_total
    pairs = []
    for name, a in zip(NAMES, got[:6]):
        ref = g[f"view_{tag}_{name}"]
        a = a.cpu().numpy()
        assert a.shape == ref.shape, name
        pairs.append((name, a, ref, 1e-3 * scale_of(ref)))
    # every pixel of every output within 1e-3 of the output's range, except a counted handful of threshold-flip pixels
    assert_bounded(pairs, ref.shape[0] if ref.ndim == 2 else int(np.prod(ref.shape[:-1])), f"view {tag}")


@pytest.mark.gpu
def test_cuda_trajectory_score_matches_reference_python(apnerf, gold):
    """probablistic_uncertainty (the body of pipeline.py:666-798) on the golden trajectory: each of the four
    predictive-information terms within 1e-3 absolute (north star tolerance on entropies) and the score."""
    g, cfg = gold
    dev = "cuda:0"
    members = _members(apnerf, cfg, dev)
    log = []
    score = apnerf.probablistic_uncertainty(
        [m[0] for m in members], [m[1] for m in members], g["traj_poses"], img_w=cfg["img_w"], img_h=cfg["img_h"],
        focal=cfg["img_w"] / 2.0, near_plane=cfg["near_plane"], render_step_size=cfg["render_step_size"],
        cone_angle=cfg["cone_angle"], alpha_thre=cfg["alpha_thre"], scale=cfg["scale"], device=dev, log=log)
    assert np.abs(np.asarray(log[0], np.float64) - g["traj_terms"]).max() <= 1e-3, (log[0], g["traj_terms"])
    assert abs(float(score) - float(g["traj_score"])) <= 2e-3


# ------------------------------------------------------------------------------------------------
# GPU: the op boundary and the remaining render wrappers against tests/golden/reference_python_ops.npz
# (OccGridEstimator.sampling / nerfacc.rendering / _update and utils.py:63-780 run from /root/reference)
# ------------------------------------------------------------------------------------------------
OPS = os.path.join(os.path.dirname(GOLDEN), "reference_python_ops.npz")


def _patterns():
    import importlib.util

    spec = importlib.util.spec_from_file_location("golden_patterns", os.path.join(os.path.dirname(GOLDEN), "patterns.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ops_estimator(apnerf, cfg, levels, seed, dev):
    from apnerf import synthetic

    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=cfg["grid_res"], levels=levels)
    b = synthetic.make_occupancy(cfg["grid_res"], seed=seed)
    if levels > 1:
        b = torch.cat([b] + [synthetic.make_occupancy(cfg["grid_res"], n_boxes=12, seed=seed + l) for l in range(1, levels)])
    est.binaries = b
    est.occs = b.flatten().float()
    return est.to(dev).eval()


@pytest.mark.gpu
def test_cuda_sampling_and_rendering_match_reference_python(apnerf, gold):
    """OccGridEstimator.sampling (occ_grid.py:95-238): ray_indices / t_starts / t_ends BIT-EXACT, with the
    density pre-filter (visibility mask from render_visibility_from_density) and on two grid levels;
    nerfacc.rendering (volrend.py:15-120) on those samples within 1e-5 relative."""
    _, cfg = gold
    g = np.load(OPS)
    P = _patterns()
    dev = "cuda:0"
    ro, rd = torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev)
    e1 = _ops_estimator(apnerf, cfg, 1, 1, dev)
    ri, ts, te = e1.sampling(ro, rd, sigma_fn=P.sigma_pattern, near_plane=0.1, far_plane=1e10, render_step_size=5e-3,
                             stratified=False, cone_angle=0.004, alpha_thre=0.01)
    assert np.array_equal(ri.cpu().numpy(), g["s1_ray_indices"])
    assert np.array_equal(ts.cpu().numpy().view(np.int32), g["s1_t_starts"].view(np.int32))
    assert np.array_equal(te.cpu().numpy().view(np.int32), g["s1_t_ends"].view(np.int32))
    e2 = _ops_estimator(apnerf, cfg, 2, 4, dev)
    ri2, ts2, te2 = e2.sampling(ro, rd, near_plane=0.2, far_plane=30.0, render_step_size=2e-2, stratified=False,
                                cone_angle=0.0, alpha_thre=0.0, early_stop_eps=0.0)
    assert np.array_equal(ri2.cpu().numpy(), g["s2_ray_indices"])
    assert np.array_equal(ts2.cpu().numpy().view(np.int32), g["s2_t_starts"].view(np.int32))
    assert np.array_equal(te2.cpu().numpy().view(np.int32), g["s2_t_ends"].view(np.int32))
    rgb, opa, dep, extras = apnerf.nerfacc.rendering(ts, te, ri, n_rays=ro.shape[0], rgb_sigma_fn=P.rgb_sigma_pattern,
                                                     render_bkgd=torch.tensor([0.1, 0.2, 0.3], device=dev))
    for name, a in (("rgb", rgb), ("opacity", opa), ("depth", dep), ("weights", extras["weights"])):
        ref = g[f"r_{name}"]
        assert np.abs(a.cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), name


@pytest.mark.gpu
def test_cuda_render_wrappers_match_reference_python(apnerf, gold):
    """render_image_with_occgrid_with_depth_guide / render_image_with_occgrid / render_image_with_occgrid_test
    (utils.py:63-780) in eval mode.  The field is fp16 (1e-3 absolute on its outputs), and its density also
    drives the visibility filter, so a few samples on the alpha threshold differ: 2 % on the sample count; every
    pixel within 1e-3 of the output's range except a counted handful of threshold-flip pixels (tests/bounds.py)."""
    from apnerf import synthetic

    _, cfg = gold
    g = np.load(OPS)
    dev = "cuda:0"
    w, h = 32, 24
    rays = apnerf.Rays(origins=torch.from_numpy(g["rays_o"]).to(dev), viewdirs=torch.from_numpy(g["rays_d"]).to(dev))
    e1 = _ops_estimator(apnerf, cfg, 1, 1, dev)

    def field(seed, C):
        f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=C)
        synthetic.init_trained_like(f, seed=seed, density_gain=cfg["density_gain"])
        return f.to(dev).eval()

    opts = dict(near_plane=cfg["near_plane"], render_step_size=4e-3, cone_angle=cfg["cone_angle"],
                alpha_thre=cfg["alpha_thre"], render_bkgd=torch.tensor([0.1, 0.2, 0.3], device=dev))

    def check(prefix, got, names):
        n_ref = int(g[f"{prefix}_n"])
        assert abs(int(got[-1]) - n_ref) <= 0.02 * n_ref, (prefix, int(got[-1]), n_ref)
        pairs = []
        for name, a in zip(names, got):
            ref = g[f"{prefix}_{name}"]
            a = a.cpu().numpy()
            assert a.shape == ref.shape, (prefix, name)
            pairs.append((name, a, ref, 1e-3 * scale_of(ref)))
        assert_bounded(pairs, w * h, prefix)

    with torch.no_grad():
        check("guide", apnerf.render_image_with_occgrid_with_depth_guide(
            field(cfg["field_seeds"][0], cfg["n_classes"]), e1, rays, depth=torch.full((w * h,), 2.0, device=dev), **opts),
            ("rgb", "opacity", "depth", "sem"))
        plain = field(cfg["field_seeds"][1], 0)
        check("occgrid", apnerf.render_image_with_occgrid(plain, e1, rays, test_chunk_size=300, **opts),
              ("rgb", "opacity", "depth"))
        check("test", apnerf.render_image_with_occgrid_test(1024, plain, e1, rays, **opts), ("rgb", "opacity", "depth"))
        # two occupancy-grid levels: the drop-in routes to the op-by-op CUDA path (sorted interval ends)
        e2 = _ops_estimator(apnerf, cfg, 2, 4, dev)
        check("lvl2", apnerf.render_probablistic_image_with_occgrid_test(
            256, field(cfg["field_seeds"][0], cfg["n_classes"]), e2, rays, near_plane=0.2, render_step_size=1e-2,
            cone_angle=0.004, alpha_thre=0.01, render_bkgd=opts["render_bkgd"]),
            ("rgb", "rgb_var", "opacity", "depth", "depth_var", "sem"))


@pytest.mark.gpu
def test_cuda_training_mode_render_matches_reference_python(apnerf, gold, monkeypatch):
    """render_image_with_occgrid_with_depth_guide with the field in TRAINING mode (utils.py:63-219: one chunk,
    stratified sampling; the differentiable field path = apnerf_field_forward_train + torch activations).  The
    stratified jitter is patched to 0.5 on both sides (tests/golden/patterns.py)."""
    from apnerf import synthetic

    _, cfg = gold
    g = np.load(OPS)
    P = _patterns()
    dev = "cuda:0"
    w, h = 32, 24
    rays = apnerf.Rays(origins=torch.from_numpy(g["rays_o"]).to(dev), viewdirs=torch.from_numpy(g["rays_d"]).to(dev))
    e1 = _ops_estimator(apnerf, cfg, 1, 1, dev)
    f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=cfg["n_classes"])
    synthetic.init_trained_like(f, seed=cfg["field_seeds"][0], density_gain=cfg["density_gain"])
    f = f.to(dev).train()
    monkeypatch.setattr(torch, "rand_like", P.half_like)
    got = apnerf.render_image_with_occgrid_with_depth_guide(
        f, e1, rays, near_plane=cfg["near_plane"], render_step_size=4e-3, cone_angle=cfg["cone_angle"],
        alpha_thre=cfg["alpha_thre"], render_bkgd=torch.tensor([0.1, 0.2, 0.3], device=dev),
        depth=torch.full((w * h,), 2.0, device=dev))
    assert got[0].requires_grad and got[3].requires_grad  # the differentiable path was taken
    n_ref = int(g["train_n"])
    assert abs(int(got[4]) - n_ref) <= 0.02 * n_ref, (int(got[4]), n_ref)
    pairs = []
    for name, a in zip(("rgb", "opacity", "depth", "sem"), got[:4]):
        ref = g[f"train_{name}"]
        a = a.detach().cpu().numpy()
        assert a.shape == ref.shape, name
        pairs.append((name, a, ref, 1e-3 * scale_of(ref)))
    assert_bounded(pairs, w * h, "train-mode render")


@pytest.mark.gpu
def test_cuda_occupancy_update_matches_reference_python(apnerf, gold, monkeypatch):
    """OccGridEstimator._update (occ_grid.py:377-437), warm-up branch, with the cell jitter patched to the cell
    centre on both sides: EMA-max values and the binarised grid are bit-exact."""
    from apnerf import synthetic

    g = np.load(OPS)
    P = _patterns()
    dev = "cuda:0"
    res = 32
    eu = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=res, levels=2).to(dev)
    eu.occs[::7] = 0.03
    eu.occs[5::11] = -1.0
    monkeypatch.setattr(torch, "rand_like", P.half_like)
    for step in (0, 16):
        eu._update(step=step, occ_eval_fn=P.occ_pattern(eu.aabbs[0], res), occ_thre=0.01, ema_decay=0.95)
    assert np.array_equal(eu.occs.cpu().numpy(), g["upd_occs"])
    assert np.array_equal(eu.binaries.cpu().numpy(), g["upd_binaries"])


@pytest.mark.gpu
def test_cuda_dataset_entry_points_match_reference_python(apnerf, gold):
    """Dataset.generate_image_rays / render_probablistic_image_from_pose / render_image_from_pose
    (habitat_to_data.py:274-549): float64 numpy images of the same shapes; fp16-MLP tolerance as above."""
    from apnerf import synthetic

    g, cfg = gold
    ops = np.load(OPS)
    dev = "cuda:0"
    members = _members(apnerf, cfg, dev)
    w, h, focal = cfg["img_w"], cfg["img_h"], cfg["img_w"] / 2.0
    # rays of one camera, then the reference's rounded-linspace subsample
    pose = torch.from_numpy(synthetic.pose_to_matrix(g["traj_poses"][3])).unsqueeze(0).float()
    K = np.array([[focal, 0, w / 2], [0, focal, h / 2], [0, 0, 1.0]])
    rs = apnerf.Dataset.generate_image_rays(pose, w, h, K, dev)
    idx = np.round(np.linspace(0, w * h - 1, 24 * 32)).astype(int)
    assert rs.origins.shape == (w * h, 3)
    assert np.array_equal(rs.origins.cpu().numpy()[idx], g["view_rays_o"])
    assert np.abs(rs.viewdirs.cpu().numpy()[idx] - g["view_rays_d"]).max() <= 2e-7

    def close(a, ref, what):
        assert a.shape == ref.shape and a.dtype == np.float64, (what, a.shape, ref.shape, a.dtype)
        n_pix = int(np.prod(ref.shape[:3]))  # [K, h, w(, D)]
        assert_bounded([(str(what), a, ref, 1e-3 * scale_of(ref))], n_pix, str(what))

    traj = g["traj_poses"]
    unc = apnerf.scoring.uncertainty_view_indices(len(traj))
    args = (w, h, focal, cfg["near_plane"], cfg["render_step_size"], cfg["scale"], cfg["cone_angle"], cfg["alpha_thre"], 4, dev)
    for m, (f, e) in enumerate(members):
        out = apnerf.Dataset.render_probablistic_image_from_pose(f, e, traj[unc], *args)
        for name, a in zip(("rgb", "rgb_var", "depth", "depth_var", "acc", "sem"), out):
            close(a, g[f"traj_m{m}_{name}"].astype(np.float64), (m, name))
    # plain variant; ops fixture: estimator seed 1 + field seed 0 == member 0, first three poses of another trajectory
    poses = ops["legacy_poses"][:3]
    out = apnerf.Dataset.render_image_from_pose(members[0][0], members[0][1], poses, *args)
    for name, a in zip(("rgb", "depth", "acc", "sem"), out):
        close(a, ops[f"pose_{name}"].astype(np.float64), name)


@pytest.mark.gpu
def test_cuda_legacy_scorer_matches_reference_python(apnerf, gold):
    """trajector_uncertainty (pipeline.py:800-916) in the one configuration the reference's own code can run (a
    single member with semantic classes).  acc_inv = mean(clip(1 / (acc + 1e-4) - 1, 0, 1e4)) amplifies the fp16
    noise of nearly transparent pixels, hence 2 % relative on that term; semantic entropy term 0.05 absolute
    (= 1e-3 on the entropy x the reference's factor 50)."""
    g, cfg = gold
    ops = np.load(OPS)
    dev = "cuda:0"
    (field, est), _ = _members(apnerf, cfg, dev)
    log = []
    unc, max_idx = apnerf.trajector_uncertainty(
        [field], [est], ops["legacy_poses"], 1, img_w=cfg["img_w"], img_h=cfg["img_h"], focal=cfg["img_w"] / 2.0,
        near_plane=cfg["near_plane"], render_step_size=cfg["render_step_size"], cone_angle=cfg["cone_angle"],
        alpha_thre=cfg["alpha_thre"], scale=cfg["scale"], device=dev, log=log)
    ref = ops["legacy_terms"]
    got = np.asarray(log[0], np.float64)
    assert got.shape == ref.shape == (4, 40)
    assert np.array_equal(max_idx, ops["legacy_max_idx"])
    assert np.abs(got[0] - ref[0]).max() == 0 and np.abs(got[1] - ref[1]).max() == 0  # one member: zero variance
    assert np.abs(got[2] - ref[2]).max() <= 0.02 * np.abs(ref[2]).max(), np.abs(got[2] - ref[2]).max()
    assert np.abs(got[3] - ref[3]).max() <= 0.05, np.abs(got[3] - ref[3]).max()
    assert abs(unc - float(ops["legacy_uncertainty"])) <= 0.02 * float(ops["legacy_uncertainty"])
