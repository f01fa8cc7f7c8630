"""CPU, world_size 2, gloo: the multi-rank host logic of the scorer -- contiguous view sharding,
per-trajectory partial sums, the single all-reduce, the final normalisation -- gives the same
scores as one rank.  (The per-pixel terms come from the oracle here; the GPU kernels are
covered by the -m gpu tests.)"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pixel_terms(rgb_var, depth_var, acc, sem):
    """Per-view sums of the four per-pixel predictive-information terms (float64), from the oracle
    formulas: sums over views/pixels are linear, which is what makes view sharding exact."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O

    V = rgb_var.shape[1]
    out = np.zeros((V, 4))
    for v in range(V):
        t = O.predictive_information(rgb_var[:, v:v + 1], depth_var[:, v:v + 1], acc[:, v:v + 1], sem[:, v:v + 1])
        n = rgb_var.shape[2]
        out[v] = [t[0] * n * 3, t[1] * n, t[2] / 3 * n, t[3] / 2 * n]  # undo the means / weights -> sums
    return out


def _worker(rank, world, port, poses_n, view_traj, per_view, n_traj, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.scoring import all_reduce_partial_sums, shard_range

    lo, hi = shard_range(poses_n, rank, world)
    sums = torch.zeros((n_traj, 4), dtype=torch.float64)
    for v in range(lo, hi):
        sums[view_traj[v]] += torch.from_numpy(per_view[v])
    sums = all_reduce_partial_sums(sums)
    if rank == 0:
        ret.put(sums.numpy())
    dist.destroy_process_group()


def test_view_sharding_world2_matches_single_rank():
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.scoring import PredictiveInformationScorer, shard_range

    rng = np.random.default_rng(0)
    E, V, P, C, T = 2, 7, 50, 29, 3
    rgb_var = rng.random((E, V, P, 3)) * 0.02
    depth_var = rng.random((E, V, P)) * 0.3
    acc = rng.random((E, V, P))
    sem = rng.standard_normal((E, V, P, C)) * 2
    view_traj = np.array([0, 0, 0, 1, 1, 2, 2])
    per_view = _pixel_terms(rgb_var, depth_var, acc, sem)
    # ranges are contiguous, balanced and cover everything
    assert [shard_range(7, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert [shard_range(256, r, 8) for r in range(8)] == [(32 * r, 32 * r + 32) for r in range(8)]
    assert [shard_range(3, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, V, view_traj, per_view, T, ret)) for r in range(2)]
    [p.start() for p in procs]
    sums = ret.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    counts = np.bincount(view_traj, minlength=T) * P
    terms = PredictiveInformationScorer.finish(sums, counts)
    from oracle import oracle as O

    for t in range(T):
        sel = view_traj == t
        ref = O.predictive_information(rgb_var[:, sel], depth_var[:, sel], acc[:, sel], sem[:, sel])
        assert np.allclose(terms[t], ref, rtol=1e-10, atol=1e-12), (t, terms[t], ref)


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.training import allreduce_gradients

    m = torch.nn.Linear(4, 3)
    with torch.no_grad():
        m.weight.fill_(1.0), m.bias.fill_(0.0)
    m.weight.grad = torch.full_like(m.weight, float(rank + 1))  # rank 0 -> 1, rank 1 -> 2
    allreduce_gradients(m)  # bias.grad is None on purpose: must become zeros, not crash
    if rank == 0:
        ret.put((m.weight.grad.clone().numpy(), m.bias.grad.clone().numpy()))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_averages():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    wg, bg = ret.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert np.allclose(wg, 1.5) and np.allclose(bg, 0.0)


def _grad_worker_empty_rank(rank, world, port, ret):
    """a rank whose batch produced no samples must still join the all-reduce (zeros,
    contributed=False) and the mean must run over the contributing ranks only."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.training import allreduce_gradients

    m = torch.nn.Linear(4, 3)
    if rank == 0:
        m.weight.grad = torch.full_like(m.weight, 3.0)
        m.bias.grad = torch.full_like(m.bias, 5.0)
    n = allreduce_gradients(m, contributed=(rank == 0))  # rank 1: no gradients at all
    n_none = allreduce_gradients(torch.nn.Linear(2, 2), contributed=False)  # nobody contributed -> 0, zero grads
    ret.put((rank, float(n), float(n_none), m.weight.grad.clone().numpy(), m.bias.grad.clone().numpy()))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_rank_without_samples():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker_empty_rank, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    got = [ret.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, n, n_none, wg, bg in got:
        assert float(n) == 1 and float(n_none) == 0
        assert np.allclose(wg, 3.0) and np.allclose(bg, 5.0), (rank, wg, bg)  # sum over ranks / 1 contributing rank


def test_ticket_counter_hands_out_disjoint_batches():
    """The scheduler's counter (one process: a local integer; several ranks: an atomic key of the process group's
    store, exercised by tests/test_multigpu_gpu.py): consecutive draws are disjoint and cover [0, n)."""
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.scoring import _Tickets

    t = _Tickets(10, None, "")
    got = [t.take(4) for _ in range(4)]
    assert got == [0, 4, 8, 12] and t.world == 1


def _ticket_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.scoring import _Tickets

    t = _Tickets(64, dist.distributed_c10d._get_default_store(), "apnerf/tickets/test/1")
    mine = []
    n_batches = 22  # batch b = every 22nd view, as the scorer deals the heavy-first order out
    while True:
        b = t.take(1)
        if b >= n_batches:
            break
        mine += list(range(64))[b::n_batches]
    dist.barrier()
    ret.put((rank, mine, t.world))
    dist.destroy_process_group()


def test_ticket_counter_world2_partitions_the_views():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_ticket_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    got = [ret.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    views = sorted(got[0][1] + got[1][1])
    assert views == list(range(64)) and got[0][2] == 2  # disjoint draws that cover every view exactly once


def test_lpt_assignment_is_balanced_and_deterministic():
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.scoring import lpt_assign

    rng = np.random.default_rng(1)
    cost = rng.uniform(0.7, 5.8, 256)  # the spread of per-view sample counts on the synthetic scene
    bins = lpt_assign(cost, 8)
    assert sorted(np.concatenate(bins).tolist()) == list(range(256))  # a partition
    loads = np.array([cost[b].sum() for b in bins])
    assert loads.max() / loads.mean() < 1.01
    contiguous = np.array([cost[32 * r:32 * r + 32].sum() for r in range(8)])
    assert loads.max() < contiguous.max()
    again = lpt_assign(cost.copy(), 8)
    assert all(np.array_equal(a, b) for a, b in zip(bins, again))
    assert [len(b) for b in lpt_assign(np.ones(5), 2)] == [3, 2] and lpt_assign(np.zeros(0), 3)[0].size == 0


def test_plan_batches_partitions_and_isolates_the_heaviest_view():
    """The renderer passes of a rank: a partition of its views; without costs even strided passes of at most
    views_per_batch views; with costs (heaviest-first order) passes of about equal cost, so a view 20x heavier than the
    median gets the FIRST pass to itself."""
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.scoring import PredictiveInformationScorer

    class Stub:  # plan_batches only reads these two attributes
        views_per_batch, min_batches, shared_passes_per_rank = 64, 3, 8

    plan = lambda order, cost, ranks=1: PredictiveInformationScorer.plan_batches(Stub, np.asarray(order), cost, ranks)
    even = plan(np.arange(72), None)
    assert [len(b) for b in even] == [36, 36] and sorted(np.concatenate(even).tolist()) == list(range(72))
    assert plan(np.arange(0), None) == []
    rng = np.random.default_rng(2)
    cost = rng.uniform(1.5, 6.0, 32)
    cost[7] = 68.0  # the camera inside a transparent box
    order = np.argsort(-cost, kind="stable")
    b = plan(order, cost)
    assert sorted(np.concatenate(b).tolist()) == list(range(32)) and len(b) >= 3
    assert b[0].tolist() == [7]  # the monster alone, first
    rest = [cost[x].sum() for x in b[1:]]
    assert max(rest) / min(rest) < 1.5
    # many views: never more than views_per_batch per pass
    cost = rng.uniform(1.0, 2.0, 300)
    b = plan(np.argsort(-cost, kind="stable"), cost)
    assert max(len(x) for x in b) <= 64 and sorted(np.concatenate(b).tolist()) == list(range(300))
    # shared counter: about four passes per rank
    b = plan(np.argsort(-cost, kind="stable"), cost, 8)
    assert 64 <= len(b) <= 72 and sorted(np.concatenate(b).tolist()) == list(range(300))


def test_draw_throttle_holds_back_the_rank_with_the_heavy_pass():
    """The shared pass queue's throttle: a rank whose drawn cost is ahead of the ranks' average does not draw while it
    has a pass in flight; with nothing in flight it always draws (no idling, no dead-lock)."""
    sys.path.insert(0, ROOT)
    import apnerf
    from apnerf.scoring import _PassQueue, _Tickets

    class Counter(_Tickets):  # a local counter that pretends to be shared by 4 ranks
        def __init__(self, n):
            super().__init__(n, None, "")
            self.world = 4

    costs = [90.0] + [10.0] * 19
    q = _PassQueue("cpu")
    q.set_shared([np.array([i]) for i in range(20)], Counter(20), costs)
    # the heavy pass (90 > 2 x mean 14) counts 1.3 x 90 = 117 and takes away this rank's slack
    assert q.may_draw(False) and int(q.next()[0]) == 0 and q.my_load == 117.0
    assert not q.may_draw(True)  # 117 > 117 / 4
    q.tickets.next = 12  # the other ranks have drawn passes 1..11 meanwhile: average (117 + 110) / 4 = 57, still behind
    assert not any(q.may_draw(True) for _ in range(12))
    q.tickets.next = 20  # everything handed out: average 77 -> still blocked, but nothing is left anyway
    assert q.may_draw(False)  # nothing in flight: always allowed
    light = _PassQueue("cpu")
    light.set_shared([np.array([i]) for i in range(20)], Counter(20), costs)
    light.tickets.next = 1
    assert int(light.next()[0]) == 1 and light.my_load == 10.0
    assert light.may_draw(True)  # 10 <= (117 + 10) / 4 + 14
