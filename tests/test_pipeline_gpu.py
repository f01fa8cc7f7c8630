"""GPU: the host glue of the planning / training loop (active-perception-..._b200/pipeline.py) and the reference's
on-disk formats -- SURVEY.md section 8 rows (f3), (f4) and BASELINE.json configs[4]:

* checkpoint ``{"occ_grid", "model", "optimizer_state_dict"}`` (scripts/pipeline.py:630-634, 1262-1274) round trip:
  a fresh mapper that loads it renders bit-identical images and continues training from the same optimizer state;
* ``uncertainty.npy`` (pipeline.py:783-790, 1256-1257): ``[planning_step, num_traj, 4]`` with the four logged terms;
* one planning round (score N trajectories -> argmax -> retrain, pipeline.py:1077-1085, 1211) end to end;
* the visualisation renders of ``ActiveNeRFMapper.render`` (pipeline.py:955-1021) and their 8-bit images.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(img_w=160, img_h=120, hfov=np.pi / 2, near_plane=0.1, render_step_size=1e-3, cone_angle=0.004,
           alpha_thre=0.01, planning_step=2, num_traj=3, cuda=DEV)


def _mapper(apnerf, seeds=(2, 12), density_gain=2.0):
    from apnerf import synthetic

    fields, ests, opts = [], [], []
    for s in seeds:
        est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
        est.binaries = synthetic.make_occupancy(128, seed=1)
        est.occs = est.binaries.flatten().float() * 0.5
        ests.append(est.to(DEV))
        f = apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=29)
        fields.append(synthetic.init_trained_like(f, seed=s, density_gain=density_gain).to(DEV))
        opts.append(torch.optim.Adam(fields[-1].parameters(), lr=1e-3, eps=1e-15))
    return apnerf.ActiveNeRFMapper(fields, ests, opts, CFG)


def _data(apnerf):
    from apnerf import synthetic

    ts = synthetic.TrainingSet(n_images=4, width=160, height=120, focal=80.0, n_classes=29, seed=4, device=DEV)
    return lambda model_idx: ts.fetch(2048)


def test_checkpoint_round_trip(apnerf, tmp_path):
    from apnerf import synthetic

    a = _mapper(apnerf)
    fetch = _data(apnerf)
    a.nerf_training(3, fetch, planning_step=1)
    path = os.path.join(tmp_path, "checkpoints", "model_0.pth")
    a.save_checkpoint(0, path)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert sorted(ck) == ["model", "occ_grid", "optimizer_state_dict"]
    assert list(ck["model"]) == ["aabb", "direction_encoding.params", "mlp_base.params", "mlp_head.params", "mlp_sem.params"]
    assert ck["occ_grid"].dtype == torch.bool and tuple(ck["occ_grid"].shape) == (1, 128, 128, 128)
    b = _mapper(apnerf, seeds=(99, 98))  # different weights, fresh optimizer
    b.load_checkpoint(0, path)
    for k, v in a.radiance_fields[0].state_dict().items():
        assert torch.equal(v, b.radiance_fields[0].state_dict()[k]), k
    assert torch.equal(a.estimators[0].binaries, b.estimators[0].binaries)
    sa, sb = a.optimizers[0].state_dict()["state"], b.optimizers[0].state_dict()["state"]
    assert sa.keys() == sb.keys() and len(sa) > 0
    for i in sa:
        for k in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(sa[i][k], sb[i][k])
    pose = synthetic.make_poses_corridor(1, seed=5)
    for m in (a, b):
        [f.eval() for f in m.radiance_fields]
        [e.eval() for e in m.estimators]
    ra, rb = a.render(pose), b.render(pose)
    for k in ("rgb_predictions", "depth_predictions", "acc_predictions", "sem_predictions"):
        assert np.array_equal(ra[k], rb[k]), k
    assert ra["rgb_predictions"].shape == (1, 120, 160, 3) and ra["pd_sem"].shape == (1, 120, 160)
    assert ra["pd_rgb"].dtype == np.float32 and ra["pd_rgb"].max() <= 255.0 and ra["pd_occ"].max() <= 255.0
    assert np.array_equal(ra["pd_dep"], np.clip(ra["depth_predictions"] * 25, 0, 255))


def test_uncertainty_log_and_planning_round(apnerf, tmp_path):
    from apnerf import synthetic

    m = _mapper(apnerf, density_gain=6.0)
    fetch = _data(apnerf)
    trajs = [synthetic.make_poses_corridor(45, seed=30 + t) for t in range(3)]
    best1, unc1 = m.planning_round(trajs, 1, 2, fetch, scale=0.25)
    assert best1 == int(np.argmax(unc1)) and unc1.shape == (3,) and np.isfinite(unc1).all()
    # the single-trajectory drop-in agrees with the batched call (before more training changes the model)
    m2 = _mapper(apnerf, density_gain=6.0)
    for mod in m2.radiance_fields + m2.estimators:
        mod.eval()
    u_batched, _ = m2.score_trajectories(trajs, 1, scale=0.25)
    m3 = _mapper(apnerf, density_gain=6.0)
    for mod in m3.radiance_fields + m3.estimators:
        mod.eval()
    u_single = [float(m3.score_trajectories([t], 1, scale=0.25)[0][0]) for t in trajs]
    assert np.abs(np.asarray(u_single) - u_batched).max() <= 1e-9
    best2, unc2 = m.planning_round(trajs, 2, 2, fetch, scale=0.25)
    m.save_all(str(tmp_path))
    log = np.load(os.path.join(tmp_path, "uncertainty.npy"))
    assert log.shape == (2, 3, 4)  # [planning_step, num_traj, (rgb, depth, 3 sem, 2 occ)]
    assert np.allclose(log[0].sum(1), unc1, rtol=0, atol=1e-12) and np.allclose(log[1].sum(1), unc2, rtol=0, atol=1e-12)
    # the reference's stop criterion reads the log this way (pipeline.py:1213-1215)
    past = np.array(m.trajector_uncertainty_list[:2]).astype(float)
    assert np.max(np.mean(past, axis=2), axis=1).shape == (2,)
    for i in range(2):
        assert os.path.exists(os.path.join(tmp_path, "checkpoints", f"model_{i}.pth"))
