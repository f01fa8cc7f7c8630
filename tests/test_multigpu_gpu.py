"""GPU, 2 ranks over NCCL: the view-sharded scorer gives the 1-rank scores (missing #10).

Every rank holds both ensemble members and the whole pose batch; the views are drawn in batches, heaviest first, from a
counter in the process group's store (``PredictiveInformationScorer.schedule``), each rank renders + scores what it drew
and ONE all-reduce of [n_traj, 4] float64 sums finishes the job.  Renders are deterministic per view and the
per-trajectory sums are float64, so the result must match the single-rank one to round-off -- with the static balanced
split ("lpt", the default), the shared counter ("dynamic") and the contiguous shard.  Skipped on a box with one GPU (run it with ``gpurun --gpus 2``)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPTS = dict(near_plane=0.1, render_step_size=1e-3, cone_angle=0.004, alpha_thre=0.01)
W, H = 80, 60


def _build(dev, balance):
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf import synthetic

    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
    est.binaries = synthetic.make_occupancy(128, seed=1)
    est = est.to(dev).eval()
    fields = [synthetic.init_trained_like(apnerf.NGPRadianceField(synthetic.ROI_AABB, layers=2, num_semantic_classes=29),
                                          seed=s).to(dev).eval() for s in (2, 12)]
    scorer = apnerf.PredictiveInformationScorer(fields, [est, est], W, H, W / 2.0, device=dev, views_per_batch=8,
                                                balance=balance, **OPTS)
    poses = synthetic.make_poses(22, seed=3)
    traj = (np.arange(22) // 8).astype(np.int32)
    return scorer, poses, traj


def _worker(rank, world, port, balance, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    scorer, poses, traj = _build(dev, balance)
    terms = scorer.score_views(poses, traj, 3)
    ret.put((rank, terms, scorer.views_rendered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("balance", ["lpt", "dynamic", "contiguous"])
def test_two_rank_scores_equal_one_rank(balance):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    scorer, poses, traj = _build(torch.device("cuda", 0), balance)
    ref = scorer.score_views(poses, traj, 3)
    del scorer
    torch.cuda.empty_cache()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, balance, ret)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted([ret.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert got[0][2] + got[1][2] == 22, "the ranks' draws must partition the batch"
    if balance == "contiguous":
        assert got[0][2] == 11 and got[1][2] == 11
    for rank, terms, _ in got:
        assert np.abs(terms - ref).max() <= 1e-9, (rank, np.abs(terms - ref).max())
