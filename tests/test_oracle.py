"""CPU: pin the oracle against the reference's golden vectors / known answers."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_weights_golden_vectors(oracle):
    # perception/nerfacc/tests/test_rendering.py:110-133 (test_grads forward values)
    ray_indices = np.array([0, 2, 2, 2, 2])
    packed_info = np.array([[0, 1], [1, 0], [1, 4]])
    sigmas = np.array([0.4, 0.8, 0.1, 0.8, 0.1], np.float32)
    t_starts = np.random.default_rng(0).random(5).astype(np.float32)
    t_ends = t_starts + 1.0
    ref = np.array([0.3297, 0.5507, 0.0428, 0.2239, 0.0174], np.float32)
    w1, _, _ = oracle.render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=3)
    w2, _, _ = oracle.render_weight_from_density(t_starts, t_ends, sigmas, packed_info=packed_info)
    assert np.allclose(w1, ref, atol=1e-4) and np.allclose(w2, ref, atol=1e-4)


def test_transmittance_docstring_example(oracle):
    # perception/nerfacc/nerfacc/volrend.py:249-256, 350-358
    t_starts = np.arange(7, dtype=np.float32)
    t_ends = t_starts + 1
    sigmas = np.array([0.4, 0.8, 0.1, 0.8, 0.1, 0.0, 0.9], np.float32)
    ray_indices = np.array([0, 0, 0, 1, 1, 2, 2])
    w, trans, alphas = oracle.render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=3)
    assert np.allclose(trans, [1.00, 0.67, 0.30, 1.00, 0.45, 1.00, 1.00], atol=6e-3)
    assert np.allclose(alphas, [0.33, 0.55, 0.095, 0.55, 0.095, 0.00, 0.59], atol=6e-3)
    assert np.allclose(w, [0.33, 0.37, 0.03, 0.55, 0.04, 0.00, 0.59], atol=6e-3)
    vis = oracle.render_visibility_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=3,
                                                early_stop_eps=0.3, alpha_thre=0.2)
    assert vis.tolist() == [True, True, False, True, False, False, True]  # volrend.py:474-476


def test_pack_info_and_scan_known_answers(oracle):
    # tests/test_pack.py:10-17, scan.py:37-40, 77-80
    assert oracle.pack_info(np.array([0, 2, 2, 2, 2]), 3).tolist() == [[0, 1], [1, 0], [1, 4]]
    assert oracle.pack_info(np.array([0, 0, 1, 1, 1, 2, 2, 2, 2]), 3).tolist() == [[0, 2], [2, 3], [5, 4]]
    x = np.arange(1, 10, dtype=np.float32)
    pi = np.array([[0, 2], [2, 3], [5, 4]])
    assert oracle.exclusive_sum(x, pi).tolist() == [0, 1, 0, 3, 7, 0, 6, 13, 21]
    # == torch.cumsum on 5 x 1000 (tests/test_scan.py:38-64)
    data = np.random.default_rng(42).random((5, 1000)).astype(np.float32)
    pi = np.stack([np.arange(5) * 1000, np.full(5, 1000)], -1)
    ex = oracle.exclusive_sum(data.reshape(-1), pi).reshape(5, 1000)
    ref = np.cumsum(np.concatenate([np.zeros((5, 1)), data[:, :-1]], 1), 1)
    assert np.allclose(ex, ref, atol=3e-4)


def test_accumulate_with_empty_ray(oracle):
    # tests/test_rendering.py:87-106
    ray_indices = np.array([0, 2, 2, 2, 2])
    weights = np.array([0.4, 0.3, 0.8, 0.8, 0.5], np.float32)
    values = np.random.default_rng(1).random((5, 2)).astype(np.float32)
    out = oracle.accumulate_along_rays(weights, values, ray_indices, 3)
    assert out.shape == (3, 2)
    assert np.allclose(out[0], weights[0] * values[0])
    assert (out[1] == 0).all()
    assert np.allclose(out[2], (weights[1:, None] * values[1:]).sum(0))


def test_ray_aabb_matches_pure_formula(oracle):
    # tests/test_grid.py:8-35 (CUDA kernel vs the reference's pure-torch _ray_aabb_intersect)
    rng = np.random.default_rng(42)
    rays_o = rng.random((1000, 3)).astype(np.float32)
    rays_d = rng.standard_normal((1000, 3)).astype(np.float32)
    rays_d /= np.linalg.norm(rays_d, axis=-1, keepdims=True)
    amin = rng.random((100, 3)).astype(np.float32)
    aabbs = np.concatenate([amin, amin + rng.random((100, 3)).astype(np.float32)], -1)
    t0, t1, h = oracle.ray_aabb_intersect(rays_o, rays_d, aabbs)
    _t0, _t1, _h = oracle._pure_ray_aabb_intersect(rays_o, rays_d, aabbs)
    assert (h == _h).all()
    assert np.allclose(t0, _t0) and np.allclose(t1, _t1)


def test_traverse_properties(oracle):
    # tests/test_grid.py:135-159: near / far planes respected to +- step / 2
    rays_o = np.array([[-1.0, 0.0, 0.0]], np.float32)
    rays_d = np.array([[1.0, 0.01, 0.01]], np.float32)
    rays_d /= np.linalg.norm(rays_d, axis=-1, keepdims=True)
    binaries = np.ones((1, 1, 1, 1), bool)
    aabbs = np.array([[0, 0, 0, 1, 1, 1]], np.float32)
    iv, sm, _ = oracle.traverse_grids(rays_o, rays_d, binaries, aabbs, np.array([1.2], np.float32),
                                      np.array([1.5], np.float32), step_size=0.05)
    assert iv["vals"].size > 0
    assert (iv["vals"] >= 1.2 - 0.025).all() and (iv["vals"] <= 1.5 + 0.025).all()
    # edges = samples + runs; left/right counts equal the sample count
    assert iv["is_left"].sum() == iv["is_right"].sum() == sm["vals"].size


def test_traverse_test_mode_consistency(oracle):
    # tests/test_grid.py:72-131: two limited over-allocated calls == one unlimited pass (atol 1e-1)
    rng = np.random.default_rng(42)
    n = 10
    rays_o = rng.standard_normal((n, 3)).astype(np.float32)
    rays_d = rng.standard_normal((n, 3)).astype(np.float32)
    rays_d /= np.linalg.norm(rays_d, axis=-1, keepdims=True)
    base = np.array([-1, -1, -1, 1, 1, 1], np.float32)

    def enlarge(a, f):
        c, e = (a[:3] + a[3:]) / 2, (a[3:] - a[:3]) / 2
        return np.concatenate([c - e * f, c + e * f])

    aabbs = np.stack([enlarge(base, 2 ** i) for i in range(4)]).astype(np.float32)
    binaries = rng.random((4, 32, 32, 32)) > 0.5
    iv, sm, _ = oracle.traverse_grids(rays_o, rays_d, binaries, aabbs)
    ts, te = iv["vals"][iv["is_left"]], iv["vals"][iv["is_right"]]
    acc_s = oracle.accumulate_along_rays(ts, None, sm["ray_indices"], n)
    acc_e = oracle.accumulate_along_rays(te, None, sm["ray_indices"], n)
    _s = _e = 0.0
    term, mask = None, None
    for _ in range(2):
        _iv, _sm, term = oracle.traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=term,
                                               traverse_steps_limit=4000, over_allocate=True, rays_mask=mask)
        mask = _sm["packed_info"][:, 1] == 4000
        ri = _sm["ray_indices"][_sm["is_valid"]]
        _s = _s + oracle.accumulate_along_rays(_iv["vals"][_iv["is_left"]], None, ri, n)
        _e = _e + oracle.accumulate_along_rays(_iv["vals"][_iv["is_right"]], None, ri, n)
    assert (~mask).all()
    assert np.allclose(_s, acc_s, atol=1e-1) and np.allclose(_e, acc_e, atol=1e-1)


def test_hashgrid_level_table(oracle):
    # SURVEY.md Appendix C level table
    meta, total = oracle.hashgrid_meta()
    assert meta[:, 1].tolist() == [16, 24, 34, 49, 71, 102, 148, 213, 308, 446, 646, 934, 1352, 1956, 2831, 4096]
    assert meta[:5, 2].tolist() == [4096, 13824, 39304, 117656, 357912]
    assert (meta[5:, 2] == 1 << 19).all()
    assert total == 6299960


def test_hashgrid_interpolation_properties(oracle):
    """On a table whose features are constant per level, interpolation returns that constant;
    at exact dense-grid nodes the encoding returns the node's own entry."""
    meta, total = oracle.hashgrid_meta()
    table = np.zeros((total, 4), np.float16)
    for l in range(16):
        table[meta[l, 3]:meta[l, 3] + meta[l, 2]] = np.float16(0.25 * (l + 1) / 4)
    x = np.random.default_rng(0).random((257, 3)).astype(np.float32)
    enc, idx = oracle.hashgrid_encode(x, table, meta, want_indices=True)
    for l in range(16):
        assert np.allclose(enc[:, 4 * l:4 * l + 4].astype(np.float32), 0.25 * (l + 1) / 4, atol=2e-3)
        assert (idx[:, l] >= meta[l, 3]).all() and (idx[:, l] < meta[l, 3] + meta[l, 2]).all()
    # level 0 is dense with scale 15: x = (i + 0.5 - 0.5) / 15 sits on node i -> weight 1 on corner 0
    table = np.random.default_rng(1).random((total, 4)).astype(np.float16)
    node = np.array([[3, 5, 7]], np.float32) / np.float32(15.0)
    enc, idx = oracle.hashgrid_encode(node, table, meta, want_indices=True)
    # pos = 15 * x + 0.5 = i + 0.5 -> cell i, frac 0.5 on every axis: mean of the 8 corners
    corners = table[idx[0, 0]].astype(np.float32)
    assert np.allclose(enc[0, :4].astype(np.float32), corners.mean(0), atol=2e-3)
    assert idx[0, 0, 0] == 3 + 5 * 16 + 7 * 256


def test_predictive_information_sanity(oracle):
    """scripts/pipeline.py:727-781: identical ensemble members carry no predictive information
    in the semantic and occupancy terms; disagreeing members do."""
    rng = np.random.default_rng(0)
    shape = (1, 4, 8, 8)
    rv = rng.random(shape + (3,)) * 0.01
    dv = rng.random(shape) * 0.01
    acc = rng.random(shape)
    sem = rng.standard_normal(shape + (29,))
    same = oracle.predictive_information(np.stack([rv, rv]), np.stack([dv, dv]), np.stack([acc, acc]),
                                         np.stack([sem, sem]))
    assert abs(same[2]) < 1e-12 and abs(same[3]) < 1e-12
    assert abs(same[0]) < 1e-12 and abs(same[1]) < 1e-12
    sem2 = rng.standard_normal(shape + (29,))
    diff = oracle.predictive_information(np.stack([rv, rv * 4]), np.stack([dv, dv * 4]),
                                         np.stack([acc, 1 - acc]), np.stack([sem, sem2]))
    assert (diff > 0).all()
