"""Golden vectors for the hot path, produced by the REFERENCE'S OWN PYTHON run in this container.

Run (CPU only, needs /root/reference; the GPU box never runs this, it only reads the .npz files):

    python tests/golden/make_golden.py          # reference_python_path.npz : one view + a trajectory score (~90 s)
    python tests/golden/make_golden.py --ops    # reference_python_ops.npz  : op boundary, wrappers, update, scorers (~60 s)

What executes, unmodified and imported from /root/reference:

  * perception/models/utils.py            render_probablistic_image_with_occgrid_test  (lines 783-1032)
  * perception/data_proc/habitat_to_data.py  Dataset.generate_image_rays / render_probablistic_image_from_pose
                                             (lines 274-301, 413-548)
  * perception/nerfacc/nerfacc/{grid,scan,volrend,pack,data_specs}.py   (the Python around the native calls)
  * scripts/pipeline.py                   ActiveNeRFMapper.probablistic_uncertainty     (lines 666-798;
                                          the method is taken out of the file with `ast`, because importing
                                          the module needs habitat_sim / lpips / matplotlib, and is executed
                                          against a stand-in `self`)

What is stood in for, because it cannot run here (no GPU, no tinycudann):

  * the three native calls the path makes through `nerfacc.cuda` -- ray_aabb_intersect, traverse_grids,
    exclusive_sum -- are served by the C oracle (oracle/apnerf_oracle.c).  The first two are proven
    bit-identical to the reference's compiled kernels on the GPU (tests/test_gpu_reference.py); the scan is
    a sequential fp32 sum per ray (the CUDA scan's summation order is unspecified, hence the 1e-5 tolerance).
  * the radiance field (tinycudann modules) is the oracle's restatement of tiny-cuda-nn (parity unpinned,
    see oracle/oracle.py) with the seeded "trained-like" parameters of synthetic.init_trained_like.

So these fixtures pin everything the reference itself defines on this path -- ray generation, the rounded
linspace subsample, the marching schedule, masks, compositing, variances, background / depth
normalisation, the view selection of a trajectory and the four predictive-information terms -- and the
tests compare both the oracle (CPU) and the CUDA path (GPU) against them.
"""
import ast
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True  # /root/reference is read-only
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=SyntaxWarning)

from oracle import oracle as O  # noqa: E402

# scene / camera configuration shared with the tests (tests/test_golden.py reads it from the npz)
CFG = dict(img_w=160, img_h=120, scale=0.1, near_plane=0.1, render_step_size=1e-3, cone_angle=0.004,
           alpha_thre=0.01, n_ensembles=2, n_classes=29, grid_res=128, occ_seeds=(1, 5), field_seeds=(2, 7),
           density_gain=6.0, traj_len=24, pose_seed=11)


# ------------------------------------------------------------------------------------------------
# stand-ins for the native module (nerfacc/cuda/csrc/nerfacc.cpp:100-129)
# ------------------------------------------------------------------------------------------------
def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t if dtype is None else t.to(dtype)


class _SaysCuda(torch.Tensor):
    """pack.py:39 refuses CPU tensors although its body (lines 40-47) is plain torch: the sample ray indices
    carry this subclass so that the reference's own lines run here.  `.device` stays the CPU."""

    @property
    def is_cuda(self):
        return True


def native_ray_aabb_intersect(rays_o, rays_d, aabbs, near_plane, far_plane, miss_value):
    t0, t1, hits = O.ray_aabb_intersect(rays_o.numpy(), rays_d.numpy(), aabbs.numpy(), near_plane, far_plane, miss_value)
    return _t(t0), _t(t1), _t(hits.astype(bool))


def native_traverse_grids(rays_o, rays_d, rays_mask, binaries, aabbs, t_sorted, t_indices, hits, near_planes,
                          far_planes, step_size, cone_angle, compute_intervals, compute_samples,
                          compute_terminate_planes, traverse_steps_limit, over_allocate):
    iv, sm, term = O.traverse_grids(
        rays_o.numpy(), rays_d.numpy(), binaries.numpy(), aabbs.numpy(), near_planes.numpy(), far_planes.numpy(),
        step_size, cone_angle, None if traverse_steps_limit <= 0 else traverse_steps_limit, over_allocate,
        rays_mask.numpy(), t_sorted.numpy(), t_indices.numpy(), hits.numpy())

    def spec(d):
        s = types.SimpleNamespace(vals=_t(d["vals"]), chunk_starts=_t(d["chunk_starts"]), chunk_cnts=_t(d["chunk_cnts"]),
                                  ray_indices=_t(d["ray_indices"]), is_left=None, is_right=None, is_valid=None)
        for k in ("is_left", "is_right", "is_valid"):
            if k in d:
                setattr(s, k, _t(d[k].astype(bool)))
        s.ray_indices = s.ray_indices.as_subclass(_SaysCuda)
        return s

    return spec(iv), spec(sm), _t(term)


def native_exclusive_sum(chunk_starts, chunk_cnts, inputs, normalize, backward):
    assert not normalize
    packed = np.stack([chunk_starts.numpy(), chunk_cnts.numpy()], -1)
    return _t(O.exclusive_sum(inputs.numpy(), packed, backward))


def import_reference():
    os.chdir(REF)  # utils.py appends relative paths to sys.path
    sys.path[:0] = [f"{REF}/perception/models", f"{REF}/perception/nerfacc", f"{REF}/perception/nerfacc/nerfacc",
                    f"{REF}/perception/data_proc"]
    for missing in ("imageio", "matplotlib", "matplotlib.pyplot", "skimage"):  # unused on this path
        if missing not in sys.modules:
            try:
                __import__(missing)
            except Exception:
                m = types.ModuleType(missing)
                m.pyplot = m.io = m.color = m
                sys.modules[missing] = m
    import nerfacc.cuda as C
    C.ray_aabb_intersect = native_ray_aabb_intersect
    C.traverse_grids = native_traverse_grids
    C.exclusive_sum = native_exclusive_sum
    # scan.py:12 does a top-level `import cuda as _C` (the reference runs with nerfacc/ itself on sys.path);
    # in this image that name would resolve to the unrelated cuda-python package
    sys.modules["cuda"] = C
    import habitat_to_data
    import utils
    from nerfacc.estimators.occ_grid import OccGridEstimator
    for mod in list(sys.modules.values()):  # modules that bound `_C` before the alias existed
        if getattr(mod, "__file__", None) and str(mod.__file__).startswith(REF) and hasattr(mod, "_C"):
            mod._C = C
    return utils, habitat_to_data.Dataset, OccGridEstimator


def reference_method(path, cls, name):
    """Source of `cls.name` cut out of a reference file with ast (no import of the module)."""
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == name:
                    return ast.get_source_segment(src, fn)
    raise KeyError(name)


# ------------------------------------------------------------------------------------------------
# the scene: two ensemble members (occupancy grid + field), built exactly like the tests do
# ------------------------------------------------------------------------------------------------
class OracleField(torch.nn.Module):
    """Stand-in for NGPRadianceField.forward (ngp.py:222-233): (rgb, density [N,1], semantics)."""

    def __init__(self, apnerf, seed, n_classes, density_gain):
        super().__init__()
        from apnerf import synthetic
        f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=n_classes)
        synthetic.init_trained_like(f, seed=seed, density_gain=density_gain)
        self.num_semantic_classes = n_classes
        self._fp = O.FieldParams(f.mlp_base.params.detach().numpy(), f.mlp_head.params.detach().numpy(),
                                 f.mlp_sem.params.detach().numpy() if n_classes > 0 else None,
                                 num_semantic_classes=n_classes)
        self._aabb = f.aabb.numpy()

    def forward(self, positions, directions):
        out = O.field_forward(positions.numpy(), directions.numpy(), self._aabb, self._fp)
        return (_t(out[0]), _t(out[1]).reshape(-1, 1)) + tuple(_t(o) for o in out[2:])

    def query_density(self, positions):
        return _t(O.field_forward(positions.numpy(), None, self._aabb, self._fp, density_only=True)).reshape(-1, 1)


def build_scene(OccGridEstimator):
    import apnerf
    from apnerf import synthetic
    fields, ests = [], []
    for occ_seed, field_seed in zip(CFG["occ_seeds"], CFG["field_seeds"]):
        est = OccGridEstimator(torch.tensor(synthetic.ROI_AABB), resolution=CFG["grid_res"], levels=1)
        est.binaries = synthetic.make_occupancy(CFG["grid_res"], seed=occ_seed)
        ests.append(est)
        fields.append(OracleField(apnerf, field_seed, CFG["n_classes"], CFG["density_gain"]))
    return fields, ests


def main():
    torch.manual_seed(0)
    utils, Dataset, OccGridEstimator = import_reference()
    from apnerf import synthetic
    fields, ests = build_scene(OccGridEstimator)
    traj = synthetic.make_poses_corridor(CFG["traj_len"], seed=CFG["pose_seed"])
    focal = CFG["img_w"] / 2.0

    # ---- fixture 1: one view through utils.render_probablistic_image_with_occgrid_test, two option sets
    from datasets.utils import Rays
    out = {}
    pose = torch.from_numpy(synthetic.pose_to_matrix(traj[3])).unsqueeze(0).float()
    K = np.array([[focal, 0, CFG["img_w"] / 2], [0, focal, CFG["img_h"] / 2], [0, 0, 1.0]])
    rs = Dataset.generate_image_rays(pose, CFG["img_w"], CFG["img_h"], K, "cpu")
    idx = np.round(np.linspace(0, len(rs.origins) - 1, 24 * 32)).astype(int)
    rays = Rays(origins=rs.origins[idx], viewdirs=rs.viewdirs[idx])
    out["view_rays_o"], out["view_rays_d"] = rays.origins.numpy(), rays.viewdirs.numpy()
    for tag, opts in (("a", dict(cone_angle=0.004, alpha_thre=0.01, render_bkgd=torch.zeros(3))),
                      ("b", dict(cone_angle=0.0, alpha_thre=0.0, render_bkgd=torch.tensor([0.2, 0.5, 0.9])))):
        with torch.no_grad():
            r = utils.render_probablistic_image_with_occgrid_test(
                1024 if tag == "a" else 96, fields[0], ests[0], rays, near_plane=CFG["near_plane"],
                render_step_size=CFG["render_step_size"] * (1 if tag == "a" else 8), **opts)
        for name, v in zip(("rgb", "rgb_var", "opacity", "depth", "depth_var", "sem"), r[:6]):
            out[f"view_{tag}_{name}"] = v.numpy()
        out[f"view_{tag}_total_samples"] = np.int64(r[6])
        print("view", tag, "total samples", r[6], "mean opacity %.4f" % float(r[2].mean()))

    # ---- fixture 2: a trajectory through pipeline.ActiveNeRFMapper.probablistic_uncertainty
    ns = dict(np=np, torch=torch, F=torch.nn.functional, Dataset=Dataset)
    exec(reference_method(f"{REF}/scripts/pipeline.py", "ActiveNeRFMapper", "probablistic_uncertainty"), ns)
    captured = {}
    real = Dataset.render_probablistic_image_from_pose

    def spy(*a, **k):
        res = real(*a, **k)
        captured.setdefault("renders", []).append(res)
        return res

    Dataset.render_probablistic_image_from_pose = staticmethod(spy)
    me = types.SimpleNamespace(
        config_file=dict(n_ensembles=CFG["n_ensembles"], cuda="cpu", img_w=CFG["img_w"], img_h=CFG["img_h"],
                         near_plane=CFG["near_plane"], render_step_size=CFG["render_step_size"],
                         cone_angle=CFG["cone_angle"], alpha_thre=CFG["alpha_thre"]),
        radiance_fields=fields, estimators=ests, focal=focal, trajector_uncertainty_list=[[]])
    score = ns["probablistic_uncertainty"](me, traj, 1)
    out["traj_poses"] = traj
    out["traj_score"] = np.float64(score)
    out["traj_terms"] = np.asarray(me.trajector_uncertainty_list[0][0], np.float64)
    for m, res in enumerate(captured["renders"]):
        for name, v in zip(("rgb", "rgb_var", "depth", "depth_var", "acc", "sem"), res):
            out[f"traj_m{m}_{name}"] = v.astype(np.float32)
    print("trajectory score", score, out["traj_terms"])

    out["cfg_keys"] = np.array(sorted(CFG))
    for k, v in CFG.items():
        out[f"cfg_{k}"] = np.asarray(v)
    dst = os.path.join(HERE, "reference_python_path.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


def main_sampling():
    """Second fixture: the op boundary (OccGridEstimator.sampling, rendering, _update) and the remaining
    render wrappers of utils.py, all from the reference's Python in eval mode (stratified=False)."""
    sys.path.insert(0, HERE)
    import patterns as P
    utils, Dataset, OccGridEstimator = import_reference()
    from apnerf import synthetic
    from datasets.utils import Rays
    from nerfacc.volrend import rendering
    out = {}
    traj = synthetic.make_poses_corridor(4, seed=CFG["pose_seed"])
    w, h = 32, 24
    pose = torch.from_numpy(synthetic.pose_to_matrix(traj[1])).unsqueeze(0).float()
    K = np.array([[w / 2.0, 0, w / 2], [0, w / 2.0, h / 2], [0, 0, 1.0]])
    rs = Dataset.generate_image_rays(pose, w, h, K, "cpu")
    out["rays_o"], out["rays_d"] = rs.origins.numpy(), rs.viewdirs.numpy()

    def estimator(levels, seed):
        est = OccGridEstimator(torch.tensor(synthetic.ROI_AABB), resolution=CFG["grid_res"], levels=levels)
        b = synthetic.make_occupancy(CFG["grid_res"], seed=seed)
        if levels > 1:  # outer levels: the same pattern, sparser
            b = torch.cat([b] + [synthetic.make_occupancy(CFG["grid_res"], n_boxes=12, seed=seed + l) for l in range(1, levels)])
        est.binaries = b
        est.occs = b.flatten().float()
        return est.eval()

    # ---- S1: sampling with a density pre-filter (1 level) and plain traversal (2 levels)
    e1, e2 = estimator(1, 1), estimator(2, 4)
    ri, ts, te = e1.sampling(rs.origins, rs.viewdirs, sigma_fn=P.sigma_pattern, near_plane=0.1, far_plane=1e10,
                             render_step_size=5e-3, stratified=False, cone_angle=0.004, alpha_thre=0.01)
    out["s1_ray_indices"], out["s1_t_starts"], out["s1_t_ends"] = ri.numpy(), ts.numpy(), te.numpy()
    ri2, ts2, te2 = e2.sampling(rs.origins, rs.viewdirs, near_plane=0.2, far_plane=30.0, render_step_size=2e-2,
                                stratified=False, cone_angle=0.0, alpha_thre=0.0, early_stop_eps=0.0)
    out["s2_ray_indices"], out["s2_t_starts"], out["s2_t_ends"] = ri2.numpy(), ts2.numpy(), te2.numpy()
    print("sampling:", len(ri), "samples after the visibility filter;", len(ri2), "plain on 2 levels")

    # ---- S2: nerfacc.rendering on S1's samples
    bk = torch.tensor([0.1, 0.2, 0.3])
    rgb, opa, dep, extras = rendering(ts, te, torch.as_tensor(ri), n_rays=w * h, rgb_sigma_fn=P.rgb_sigma_pattern,
                                      render_bkgd=bk)
    out["r_rgb"], out["r_opacity"], out["r_depth"] = rgb.numpy(), opa.numpy(), dep.numpy()
    out["r_weights"] = extras["weights"].numpy()

    # ---- S3: the render wrappers of utils.py, eval mode, oracle field
    field = OracleField(__import__("apnerf"), CFG["field_seeds"][0], CFG["n_classes"], CFG["density_gain"]).eval()
    plain = OracleField(__import__("apnerf"), CFG["field_seeds"][1], 0, CFG["density_gain"]).eval()
    rays = Rays(origins=rs.origins, viewdirs=rs.viewdirs)
    opts = dict(near_plane=CFG["near_plane"], render_step_size=4e-3, cone_angle=CFG["cone_angle"],
                alpha_thre=CFG["alpha_thre"], render_bkgd=bk)
    with torch.no_grad():
        g = utils.render_image_with_occgrid_with_depth_guide(field, e1, rays, depth=torch.full((w * h,), 2.0), **opts)
        for name, v in zip(("rgb", "opacity", "depth", "sem"), g[:4]):
            out[f"guide_{name}"] = v.numpy()
        out["guide_n"] = np.int64(g[4])
        g = utils.render_image_with_occgrid(plain, e1, rays, test_chunk_size=300, **opts)
        for name, v in zip(("rgb", "opacity", "depth"), g[:3]):
            out[f"occgrid_{name}"] = v.numpy()
        out["occgrid_n"] = np.int64(g[3])
        g = utils.render_image_with_occgrid_test(1024, plain, e1, rays, **opts)
        for name, v in zip(("rgb", "opacity", "depth"), g[:3]):
            out[f"test_{name}"] = v.numpy()
        out["test_n"] = np.int64(g[3])
        # two grid levels (sorted interval ends, grid.py:158-162 / utils.py:884-890) through the probabilistic renderer
        g = utils.render_probablistic_image_with_occgrid_test(256, field, e2, rays, near_plane=0.2, render_step_size=1e-2,
                                                              cone_angle=0.004, alpha_thre=0.01, render_bkgd=bk)
        for name, v in zip(("rgb", "rgb_var", "opacity", "depth", "depth_var", "sem"), g[:6]):
            out[f"lvl2_{name}"] = v.numpy()
        out["lvl2_n"] = np.int64(g[6])
    # training mode (utils.py:63-219 with radiance_field.training): one chunk, STRATIFIED sampling -- the jitter
    # torch.rand_like(near_planes) * render_step_size (occ_grid.py:158-159) is patched to the constant 0.5 here
    # and in the test, the only way to compare a random path across implementations
    field.train()
    real_rand_like = torch.rand_like
    torch.rand_like = P.half_like
    try:
        g = utils.render_image_with_occgrid_with_depth_guide(field, e1, rays, depth=torch.full((w * h,), 2.0), **opts)
    finally:
        torch.rand_like = real_rand_like
        field.eval()
    for name, v in zip(("rgb", "opacity", "depth", "sem"), g[:4]):
        out[f"train_{name}"] = v.detach().numpy()
    out["train_n"] = np.int64(g[4])
    print("wrappers: samples", int(out["guide_n"]), int(out["occgrid_n"]), int(out["test_n"]), int(out["lvl2_n"]),
          int(out["train_n"]))

    # ---- S4: OccGridEstimator._update, warm-up branch (all cells), jitter patched to the cell centre
    res = 32
    eu = OccGridEstimator(torch.tensor(synthetic.ROI_AABB), resolution=res, levels=2)
    eu.occs[::7] = 0.03
    eu.occs[5::11] = -1.0  # cells no camera sees (mark_invisible_cells) are skipped
    real_rand_like = torch.rand_like
    torch.rand_like = P.half_like
    try:
        for step in (0, 16):
            eu._update(step=step, occ_eval_fn=P.occ_pattern(eu.aabbs[0], res), occ_thre=0.01, ema_decay=0.95)
    finally:
        torch.rand_like = real_rand_like
    out["upd_occs"], out["upd_binaries"] = eu.occs.numpy(), eu.binaries.numpy()
    print("update: occupied", int(eu.binaries.sum()), "of", eu.binaries.numel())

    # ---- S5: Dataset.render_image_from_pose (habitat_to_data.py:304-411) and the older scorer
    # ActiveNeRFMapper.trajector_uncertainty (pipeline.py:800-916).  The reference's own tuple unpacking (:823 wants
    # 4 values from member 0, :842 wants 3 from the others) only works for ONE member with semantic classes, so
    # that is the configuration pinned here.
    traj = synthetic.make_poses_corridor(22, seed=CFG["pose_seed"] + 1)
    focal = CFG["img_w"] / 2.0
    res = Dataset.render_image_from_pose(field, e1, traj[:3], CFG["img_w"], CFG["img_h"], focal, CFG["near_plane"],
                                         CFG["render_step_size"], CFG["scale"], CFG["cone_angle"], CFG["alpha_thre"], 4, "cpu")
    for name, v in zip(("rgb", "depth", "acc", "sem"), res):
        out[f"pose_{name}"] = v.astype(np.float32)
    ns = dict(np=np, torch=torch, F=torch.nn.functional, Dataset=Dataset)
    exec(reference_method(f"{REF}/scripts/pipeline.py", "ActiveNeRFMapper", "trajector_uncertainty"), ns)
    me = types.SimpleNamespace(
        config_file=dict(n_ensembles=1, cuda="cpu", img_w=CFG["img_w"], img_h=CFG["img_h"], near_plane=CFG["near_plane"],
                         render_step_size=CFG["render_step_size"], cone_angle=CFG["cone_angle"], alpha_thre=CFG["alpha_thre"]),
        radiance_fields=[field], estimators=[e1], focal=focal, trajector_uncertainty_list=[[]])
    unc, max_idx = ns["trajector_uncertainty"](me, traj, 1)
    out["legacy_poses"], out["legacy_uncertainty"], out["legacy_max_idx"] = traj, np.float64(unc), max_idx
    out["legacy_terms"] = np.asarray(me.trajector_uncertainty_list[0][0], np.float64)
    print("legacy scorer:", unc, out["legacy_terms"].shape)

    dst = os.path.join(HERE, "reference_python_ops.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    if "--ops" in sys.argv:
        main_sampling()
    else:
        main()
