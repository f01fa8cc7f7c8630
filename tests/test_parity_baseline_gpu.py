"""GPU parity at the BASELINE.json sizes, with hard bounds instead of quantiles.

* whole 320x240 views (76 800-ray calls, the size the per-call schedule n = R // n_alive depends on) of the bench's
  own pose set through both ensemble members, CUDA renderer vs the CPU oracle:
    - every ray's DISCRETE decisions are compared: the number of samples marched / evaluated and the number that
      passed alpha_thre and were composited.  A ray where either total differs took a different branch at a threshold
      (alpha >= alpha_thre, opacity <= 1 - early_stop_eps) because the fp16 network outputs differ in the last bit:
      a "threshold-flip ray".  Their COUNT is asserted (<= 0.1 % of the rays);
    - on every other ray (same decisions) EVERY pixel of EVERY output is bounded: colour, opacity and the colour
      variance by 1e-3 absolute (the north star's fp16-MLP tolerance), depth / depth variance / semantic logits by
      1e-3 of the output's range (they are sums of weight x value with |value| up to the scene depth / logit range);
    - the per-view predictive-information terms agree to 1e-3 absolute (entropy tolerance of the north star).
* the 256-pose planner batch (configs[2]) through the public scorer: partition invariance (one batch of 256 ==
  four batches of 64 == per-view sums), and the oracle-rendered views scored alone match the oracle's terms;
* one scale = 1 pose of the visualisation path (640x640 = 409 600 rays, pipeline.py:955-974) through the
  ``ActiveNeRFMapper.render`` drop-in against the oracle, same bounds.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OPTS = dict(near_plane=0.1, render_step_size=1e-3, cone_angle=0.004, alpha_thre=0.01)
W, H, FOCAL, C = 320, 240, 160.0, 29
VIEWS = (0, 3)           # poses of the bench's set rendered by the oracle (about 4 s of CPU each per member)
FLIP_FRACTION = 1e-3     # threshold-flip rays allowed per render
NAMES = ("rgb", "rgb_var", "opacity", "depth", "depth_var", "sem")


def _bounds(ref):
    """Absolute bound per output: 1e-3 on colour / opacity / colour variance; 1e-3 of the range elsewhere."""
    rng = lambda a: max(1.0, float(np.abs(a).max()))
    return dict(rgb=1e-3, rgb_var=1e-3, opacity=1e-3, depth=1e-3 * rng(ref["depth"]),
                depth_var=1e-3 * rng(ref["depth_var"]), sem=1e-3 * rng(ref["sem"]))


@pytest.fixture(scope="module")
def scene(apnerf):
    from apnerf import synthetic

    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
    est.binaries = synthetic.make_occupancy(128, seed=1)
    est = est.to(DEV).eval()
    fields = []
    for s in (2, 12):
        f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=C)
        fields.append(synthetic.init_trained_like(f, seed=s, density_gain=6.0).to(DEV).eval())
    return est, fields, synthetic.make_poses(256, seed=3)


def _oracle_fn(oracle, field):
    fp = oracle.FieldParams(field.mlp_base.params.detach().cpu().numpy(), field.mlp_head.params.detach().cpu().numpy(),
                            field.mlp_sem.params.detach().cpu().numpy(), num_semantic_classes=C)
    aabb = field.aabb.cpu().numpy()
    return lambda p, d: oracle.field_forward(p, d, aabb, fp)


def _oracle_render(oracle, field, est, o, d):
    counts = {}
    r = oracle.render_probablistic_image_with_occgrid_test(
        1024, _oracle_fn(oracle, field), est.binaries.cpu().numpy(), est.aabbs.cpu().numpy(), o, d, C,
        ray_counts=counts, **OPTS)
    out = dict(zip(NAMES, r[:6]))
    out["evaluated"], out["visible"], out["total"] = counts["evaluated"], counts["visible"], r[6]
    return out


def _cuda_render(apnerf, field, est, o, d, rays_per_call):
    r = apnerf.FusedRenderer(DEV, C)
    n = o.shape[0]
    counts = torch.zeros((2, n), dtype=torch.int32, device=DEV)
    st = r.render(field, est, torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), rays_per_call,
                  max_samples=1024, ray_counts=counts, **OPTS)
    r.check_overflow()
    out = {k: v.cpu().numpy() for k, v in r.finalize(st, torch.zeros(3, device=DEV)).items()}
    out["evaluated"], out["visible"] = counts[0].cpu().numpy().astype(np.int64), counts[1].cpu().numpy().astype(np.int64)
    out["total"] = int(r.total_samples[: n // rays_per_call].sum().item())
    return out


def _compare(tag, got, ref):
    """Flip count + hard bounds on all other rays; returns the report line."""
    n = ref["opacity"].shape[0]
    same = (got["evaluated"] == ref["evaluated"]) & (got["visible"] == ref["visible"])
    flips = int((~same).sum())
    bounds = _bounds(ref)
    worst = {}
    for k in NAMES:
        err = np.abs(got[k].reshape(n, -1).astype(np.float64) - ref[k].reshape(n, -1))[same]
        worst[k] = float(err.max()) if err.size else 0.0
    line = (f"{tag}: rays {n}, evaluated samples {int(ref['evaluated'].sum())} (cuda {int(got['evaluated'].sum())}), "
            f"threshold-flip rays {flips} ({100.0 * flips / n:.4f} %), worst |err| on the other rays: "
            + ", ".join(f"{k} {worst[k]:.2e} (bound {bounds[k]:.1e})" for k in NAMES))
    print(line)
    assert flips <= FLIP_FRACTION * n, line
    for k in NAMES:
        assert worst[k] <= bounds[k], line
    return line


@pytest.fixture(scope="module")
def rendered_views(apnerf, oracle, scene):
    """Oracle and CUDA renders of the full-resolution views, shared by the tests below."""
    from apnerf import synthetic

    est, fields, poses = scene
    out = {}
    for v in VIEWS:
        o, d = oracle.generate_image_rays(synthetic.pose_to_matrix(poses[v]).astype(np.float32), W, H, FOCAL)
        for m, f in enumerate(fields):
            out[(v, m)] = (_cuda_render(apnerf, f, est, o, d, W * H), _oracle_render(oracle, f, est, o, d))
    return out


def test_full_320x240_views_hard_bounds(rendered_views):
    for (v, m), (got, ref) in rendered_views.items():
        _compare(f"view {v} member {m}", got, ref)


def test_full_view_predictive_information(apnerf, oracle, scene, rendered_views):
    """The four per-view terms through the public scorer (host poses in) against float64 scoring of the oracle's
    renders: 1e-3 absolute on each entropy term."""
    est, fields, poses = scene
    scorer = apnerf.PredictiveInformationScorer(fields, [est, est], W, H, FOCAL, device=DEV, views_per_batch=2, **OPTS)
    terms = scorer.score_views(poses[list(VIEWS)], np.arange(len(VIEWS), dtype=np.int32), len(VIEWS))
    for i, v in enumerate(VIEWS):
        refs = [rendered_views[(v, m)][1] for m in range(2)]
        stack = lambda k: np.stack([r[k] for r in refs])[:, None]
        ref = oracle.predictive_information(stack("rgb_var"), stack("depth_var")[..., 0], stack("opacity")[..., 0],
                                            stack("sem"))
        print(f"view {v}: cuda terms {terms[i].tolist()} oracle {ref.tolist()}")
        assert np.abs(terms[i] - ref).max() <= 1e-3, (v, terms[i], ref)


def test_256_pose_batch_partition_invariance(apnerf, scene):
    """configs[2] at full size: the score of the 256-pose batch does not depend on how the views are batched, and a
    trajectory's sums are the sums of its views' (the reduction is linear in the views, which is also what makes the
    multi-GPU shard exact).  Renders are deterministic per view, so the sums agree to float64 round-off."""
    est, fields, poses = scene
    traj = (np.arange(256) // 32).astype(np.int32)
    one = apnerf.PredictiveInformationScorer(fields, [est, est], W, H, FOCAL, device=DEV, views_per_batch=256, **OPTS)
    four = apnerf.PredictiveInformationScorer(fields, [est, est], W, H, FOCAL, device=DEV, views_per_batch=64, **OPTS)
    t1 = one.score_views(poses, traj, 8)
    del one
    torch.cuda.empty_cache()
    t4 = four.score_views(poses, traj, 8)
    assert np.isfinite(t1).all() and (t1[:, :2] > 0).all()
    assert np.abs(t1 - t4).max() <= 1e-9, np.abs(t1 - t4).max()
    # per-view scoring of trajectory 0 (32 single-view "trajectories") sums back to the trajectory's terms
    per_view = four.score_views(poses[:32], np.arange(32, dtype=np.int32), 32)
    assert np.abs(per_view.mean(0) - t1[0]).max() <= 1e-9


def test_scale1_visualisation_render(apnerf, oracle, scene):
    """pipeline.py:955-974: one pose at scale = 1 of the 640x640 camera (409 600 rays) through the
    ActiveNeRFMapper.render drop-in (member 0, no variances) against the oracle."""
    from apnerf import synthetic

    est, fields, poses = scene
    cfg = dict(img_w=640, img_h=640, hfov=np.pi / 2, planning_step=1, cuda=DEV, **OPTS)
    mapper = apnerf.ActiveNeRFMapper(fields[:1], [est], [None], cfg)
    pose = poses[0]
    res = mapper.render(pose[None])
    assert res["rgb_predictions"].shape == (1, 640, 640, 3) and res["pd_sem"].shape == (1, 640, 640)
    o, d = oracle.generate_image_rays(synthetic.pose_to_matrix(pose).astype(np.float32), 640, 640, mapper.focal)
    ref = _oracle_render(oracle, fields[0], est, o, d)
    got = _cuda_render(apnerf, fields[0], est, o, d, 640 * 640)
    _compare("scale=1 640x640", got, ref)
    # the drop-in's arrays are those renders (the plain renderer composites the same samples without the variances)
    n = 640 * 640
    assert np.abs(res["rgb_predictions"].reshape(n, 3) - got["rgb"]).max() <= 1e-6
    assert np.abs(res["depth_predictions"].reshape(n) - got["depth"][:, 0]).max() <= 1e-5 * max(1.0, got["depth"].max())
    assert np.abs(res["acc_predictions"].reshape(n) - got["opacity"][:, 0]).max() <= 1e-6
    assert np.array_equal(res["pd_occ"], np.clip(res["acc_predictions"] * 255, 0, 255))
