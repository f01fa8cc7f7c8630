"""GPU: the training side of the path (BASELINE.json configs[3]) -- hash-grid backward kernel against
a numpy scatter of the oracle's own cell indices, the differentiable field against the fused inference
kernel (same rounding points) and against an fp32 torch restatement for gradients, and a short
optimisation run of the reference's loss (scripts/pipeline.py:507-532)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
AABB = [-6.4, -0.2, -6.4, 6.4, 12.6, 6.4]


def _field(apnerf, seed=2, C=29):
    from apnerf import synthetic

    f = apnerf.NGPRadianceField(AABB, layers=2, num_semantic_classes=C)
    synthetic.init_trained_like(f, seed=seed, density_gain=2.0)
    return f.to(DEV)


def test_hashgrid_backward_matches_scatter(apnerf, oracle):
    from apnerf._lib import call
    from apnerf.radiance_fields.ngp import hashgrid_levels

    meta, total = hashgrid_levels(16, 16, 4096, 19)
    g = torch.Generator().manual_seed(3)
    n = 3000
    x = torch.rand((n, 3), generator=g)
    d_enc = torch.randn((n, 64), generator=g)
    d_table = torch.zeros((total, 4), device=DEV)
    call("apnerf_hashgrid_encode_bwd", n, x.to(DEV), 16, meta.ctypes.data_as(ctypes.c_void_p), d_enc.to(DEV), d_table)
    # oracle: the encoding is linear in the table -> gradient = scatter of (corner weight x d_enc)
    ometa = oracle.hashgrid_meta()[0]
    _, idx = oracle.hashgrid_encode(x.numpy(), np.zeros((total, 4), np.float16), ometa, want_indices=True)
    ref = np.zeros((total, 4), np.float64)
    xs = x.numpy()
    for l in range(16):
        scale = ometa[l, 0:1].view(np.float32)[0]
        pos = (np.float32(scale) * xs + np.float32(0.5)).astype(np.float32)
        w = pos - np.floor(pos)
        for c in range(8):
            wt = np.ones(n, np.float32)
            for a in range(3):
                wt = wt * (w[:, a] if (c >> a) & 1 else (np.float32(1.0) - w[:, a]))
            np.add.at(ref, idx[:, l, c].astype(np.int64), wt[:, None].astype(np.float64) * d_enc.numpy()[:, 4 * l:4 * l + 4])
    got = d_table.cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    assert (got != 0).sum() > 1000


def test_training_forward_matches_inference_kernel(apnerf):
    f = _field(apnerf)
    g = torch.Generator().manual_seed(0)
    lo, hi = torch.tensor(AABB[:3]), torch.tensor(AABB[3:])
    pos = (lo + (hi - lo) * torch.rand((4000, 3), generator=g)).to(DEV)
    dirs = torch.randn((4000, 3), generator=g)
    dirs = (dirs / dirs.norm(dim=-1, keepdim=True)).to(DEV)
    f.eval()
    with torch.no_grad():
        rgb0, dens0, sem0 = f(pos, dirs)
    f.train()
    rgb1, dens1, sem1 = f(pos, dirs)  # differentiable path
    assert rgb1.requires_grad and dens1.requires_grad and sem1.requires_grad
    # same operands and rounding points; only the fp32 accumulation order inside the GEMMs differs
    assert (rgb0 - rgb1).abs().max() <= 1e-3
    assert float(((dens0 - dens1.detach()).abs() / dens0.clamp_min(1e-6)).max()) <= 2e-2  # every sample: <= one fp16 ulp of the logit
    assert float((sem0 - sem1.detach()).abs().max()) <= 4e-3 * max(1.0, float(sem0.abs().max()))


def test_field_gradients_match_fp32_reference(apnerf):
    """Gradients w.r.t. the MLP weights and the hash table against a plain fp32 torch restatement of
    the same network (the fp16 rounding of the forward pass bounds the agreement: a few per cent)."""
    from apnerf._lib import call
    from apnerf.radiance_fields.ngp import hashgrid_levels

    f = _field(apnerf)
    f.train()
    g = torch.Generator().manual_seed(1)
    n = 2048
    lo, hi = torch.tensor(AABB[:3]), torch.tensor(AABB[3:])
    pos = (lo + (hi - lo) * torch.rand((n, 3), generator=g)).to(DEV)
    dirs = torch.randn((n, 3), generator=g)
    dirs = (dirs / dirs.norm(dim=-1, keepdim=True)).to(DEV)
    tgt_rgb = torch.rand((n, 3), generator=g).to(DEV)
    tgt_sem = torch.randint(0, 29, (n,), generator=g).to(DEV)

    def loss_of(rgb, dens, sem):
        return ((rgb - tgt_rgb) ** 2).mean() + (torch.log1p(dens)).mean() * 0.1 + \
            torch.nn.functional.cross_entropy(sem, tgt_sem) * 0.5

    rgb, dens, sem = f(pos, dirs)
    loss_of(rgb, dens, sem).backward()
    got = {k: p.grad.clone() for k, p in f.named_parameters() if p.grad is not None}

    # fp32 restatement: gather with the kernel's own cell indices, fp32 MLPs, torch autograd
    meta, total = hashgrid_levels(16, 16, 4096, 19)
    x = ((pos - lo.to(DEV)) / (hi - lo).to(DEV)).contiguous()
    idx = torch.empty((n, 16, 8), dtype=torch.int32, device=DEV)
    call("apnerf_hashgrid_encode", n, x, 16, meta.ctypes.data_as(ctypes.c_void_p),
         f.mlp_base.params[f._n_base_w:].detach().half().contiguous(), None, idx)
    base_p = f.mlp_base.params.detach().clone().requires_grad_(True)
    head_p = f.mlp_head.params.detach().clone().requires_grad_(True)
    sem_p = f.mlp_sem.params.detach().clone().requires_grad_(True)
    table = base_p[f._n_base_w:].view(-1, 4)
    scales = torch.from_numpy(meta[:, 0].copy().view(np.float32)).to(DEV)
    p_ = scales[None, :, None] * x[:, None, :] + 0.5
    w = p_ - torch.floor(p_)
    enc = []
    for c in range(8):
        wt = torch.ones((n, 16), device=DEV)
        for a in range(3):
            wt = wt * (w[..., a] if (c >> a) & 1 else 1 - w[..., a])
        enc.append(wt[..., None] * table[idx[:, :, c].long()])
    enc = sum(enc).reshape(n, 64)
    w1, w2, w3 = f._split(base_p[: f._n_base_w], f._base_dims)
    base = torch.relu(torch.relu(enc @ w1.t()) @ w2.t()) @ w3.t()
    dens_r = torch.exp(base[:, :1] - 1)
    feat = base[:, 1:16]
    sh = torch.empty((n, 16), device=DEV, dtype=torch.float16)
    call("apnerf_sh4", n, dirs.contiguous(), sh)
    ones = torch.ones((n, 1), device=DEV)
    wh1, wh2, wh3 = f._split(head_p, f._head_dims)
    rgb_r = torch.sigmoid((torch.relu(torch.relu(torch.cat([sh.float(), feat, ones], -1) @ wh1.t()) @ wh2.t()) @ wh3.t())[:, :3])
    ws1, ws2, ws3 = f._split(sem_p, f._sem_dims)
    sem_r = (torch.relu(torch.relu(torch.cat([feat, ones], -1) @ ws1.t()) @ ws2.t()) @ ws3.t())[:, :29]
    loss_of(rgb_r, dens_r, sem_r).backward()
    for name, ref in (("mlp_base.params", base_p.grad), ("mlp_head.params", head_p.grad), ("mlp_sem.params", sem_p.grad)):
        a, b = got[name], ref
        cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0)
        rel = (a - b).norm() / b.norm()
        # fp16 activations and gradients (x128 loss scale) against an fp32 graph: measured 0.2-1 % (tools/train_grad_check.py)
        assert cos > 0.9995 and rel < 2e-2, (name, float(cos), float(rel))
    assert got["direction_encoding.params"].numel() == 0 if "direction_encoding.params" in got else True


def test_training_steps_reduce_loss(apnerf):
    """A few optimisation steps of the reference loss on one fixed synthetic batch (coherent rays from one
    origin, like habitat_to_data.py:209-218) must reduce the loss; every parameter gets a gradient."""
    from apnerf import synthetic, training

    torch.manual_seed(0)
    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
    est.binaries = synthetic.make_occupancy(128, seed=1)
    est.occs = est.binaries.flatten().float() * 0.5  # consistent with the binaries for sampling()'s alpha_thre
    est = est.to(DEV)
    f = _field(apnerf, seed=4)
    opt = torch.optim.Adam(f.parameters(), lr=1e-3, eps=1e-15)  # pipeline.py:173-178
    g = torch.Generator().manual_seed(4)
    n = 1024
    d = torch.randn((n, 3), generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    batch = dict(rays=apnerf.Rays(origins=torch.tensor([0.1, 1.5, -0.2]).expand(n, 3).contiguous().to(DEV), viewdirs=d.to(DEV)),
                 pixels=torch.rand((n, 3), generator=g).to(DEV), dep=(torch.rand(n, generator=g) * 4 + 0.5).to(DEV),
                 sem=torch.randint(0, 29, (n,), generator=g).to(DEV), color_bkgd=torch.rand(3, generator=g).to(DEV))
    losses = []
    for step in range(40):
        torch.manual_seed(100)  # same stratified jitter every step: the loss is then a fixed function
        out = training.training_step(f, est, opt, batch, step=step + 1000, update_occupancy=False)
        assert out is not None and out["n_samples"] > 0
        losses.append(out["loss"])
    assert all(p.grad is not None for p in f.parameters())
    assert losses[-1] < 0.8 * losses[0], (losses[0], losses[-1])


def test_train_mode_render_shapes_and_eval_chunking(apnerf):
    from apnerf import synthetic

    est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=128, levels=1)
    est.binaries = synthetic.make_occupancy(128, seed=1)
    est = est.to(DEV).eval()
    f = _field(apnerf).eval()
    g = torch.Generator().manual_seed(0)
    d = torch.randn((20, 30, 3), generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    rays = apnerf.Rays(origins=torch.tensor([0.1, 1.5, -0.2]).expand(20, 30, 3).contiguous().to(DEV), viewdirs=d.to(DEV))
    with torch.no_grad():
        rgb, acc, depth, sem, n = apnerf.render_image_with_occgrid_with_depth_guide(
            f, est, rays, near_plane=0.1, render_step_size=1e-3, render_bkgd=torch.ones(3, device=DEV),
            cone_angle=0.004, alpha_thre=0.01, test_chunk_size=256)
    assert rgb.shape == (20, 30, 3) and acc.shape == (20, 30, 1) and depth.shape == (20, 30, 1)
    assert sem.shape == (20, 30, 29) and n > 0
    assert torch.isfinite(rgb).all() and (acc >= 0).all() and (acc <= 1 + 1e-5).all()


def test_fused_occupancy_update_matches_op_by_op(apnerf):
    """apnerf_occ_update (jittered cell -> density -> EMA-max in the field kernel) against the op-by-op body of
    OccGridEstimator._update (occ_grid.py:377-437) with the same random numbers: bit-identical occupancy values
    and binaries in the warm-up branch (every visible cell once), and in the sampled branch everywhere except
    cells drawn more than once (the reference's indexed assignment keeps an unspecified one of them)."""
    from apnerf import synthetic
    from apnerf.nerfacc import DensityOccEvalFn

    field = _field(apnerf)
    step_size = 5e-3

    def run(fused, steps):
        est = apnerf.OccGridEstimator(synthetic.ROI_AABB, resolution=64, levels=2).to(DEV).train()
        est.occs[3::17] = -1.0  # cells no camera sees are never touched
        fn = DensityOccEvalFn(field, step_size)
        plain = (lambda x: field.query_density(x) * step_size)
        drawn = []
        sample = est._sample_uniform_and_occupied_cells
        est._sample_uniform_and_occupied_cells = lambda n: drawn.append(sample(n)) or drawn[-1]
        torch.manual_seed(123)
        for step in steps:
            est.update_every_n_steps(step=step, occ_eval_fn=fn if fused else plain, occ_thre=1e-2)
        return est.occs.clone(), est.binaries.clone(), drawn

    with torch.no_grad():
        a_occ, a_bin, _ = run(True, (0, 16))
        b_occ, b_bin, _ = run(False, (0, 16))
    assert torch.equal(a_occ, b_occ) and torch.equal(a_bin, b_bin)
    assert 0 < int(a_bin.sum()) < a_bin.numel() and bool((a_occ[3::17] == -1.0).all())
    with torch.no_grad():
        a_occ, a_bin, a_drawn = run(True, (0, 256))
        b_occ, b_bin, b_drawn = run(False, (0, 256))
    cells = a_occ.numel() // 2
    once = torch.ones_like(a_occ, dtype=torch.bool)
    for lvl, (ia, ib) in enumerate(zip(a_drawn[0], b_drawn[0])):
        assert torch.equal(ia, ib)  # same random stream on both paths
        counts = torch.bincount(ia, minlength=cells)
        once[lvl * cells:(lvl + 1) * cells] = counts <= 1
    assert 0.5 < once.float().mean().item() < 1.0
    assert torch.equal(a_occ[once], b_occ[once])
    assert (a_bin == b_bin).float().mean().item() >= 0.98


def test_training_without_semantics_and_edge_inputs(apnerf):
    """A field without a semantic head trains (the backward kernel runs with no semantic gradient), query_density is
    differentiable, empty batches are accepted, and the saved-activation path reproduces the inference kernel exactly
    for a number of points that is not a multiple of the 128-row tile."""
    f = _field(apnerf, seed=5, C=0).train()
    g = torch.Generator().manual_seed(2)
    n = 1000  # not a multiple of 128
    lo, hi = torch.tensor(AABB[:3]), torch.tensor(AABB[3:])
    pos = (lo + (hi - lo) * torch.rand((n, 3), generator=g)).to(DEV)
    pos[:7] = torch.tensor(AABB[3:]).to(DEV) + 1.0  # outside the aabb: the selector zeroes the density and its gradient
    dirs = torch.randn((n, 3), generator=g)
    dirs = (dirs / dirs.norm(dim=-1, keepdim=True)).to(DEV)
    out = f(pos, dirs)
    assert len(out) == 2
    rgb, dens = out
    assert float(dens[:7].detach().abs().max()) == 0.0
    (rgb.square().mean() + dens.mean() * 1e-2).backward()
    grads = {k: p.grad for k, p in f.named_parameters()}
    assert set(grads) == {"direction_encoding.params", "mlp_base.params", "mlp_head.params"}
    assert all(v is not None and torch.isfinite(v).all() for v in grads.values())
    assert float(grads["mlp_base.params"].abs().sum()) > 0 and float(grads["mlp_head.params"].abs().sum()) > 0
    f.eval()
    with torch.no_grad():
        rgb0, dens0 = f(pos, dirs)
    assert torch.equal(rgb0, rgb.detach()) and torch.equal(dens0, dens.detach())  # same kernel, same rounding points
    # differentiable query_density
    f.train()
    f.zero_grad()
    d2 = f.query_density(pos)
    assert d2.shape == (n, 1) and d2.requires_grad
    d2.sum().backward()
    assert float(f.mlp_base.params.grad.abs().sum()) > 0
    # empty input
    e = f(pos[:0], dirs[:0])
    assert e[0].shape == (0, 3) and e[1].shape == (0, 1)
    (e[0].sum() + e[1].sum()).backward()


@pytest.mark.parametrize("n,C", [(1, 29), (31, 29), (4096 + 17, 29), (70001, 5), (150000, 0)])
def test_wgrad_tcgen05_matches_library_gemm(apnerf, n, C):
    """apnerf_field_wgrad (tcgen05 split-K, MN-major operands, nine accumulators in TMEM) against the same nine
    products done by a library GEMM on the same fp16 matrices.  fp16 x fp16 products are exact in fp32, so the two
    differ only by the fp32 summation order (up to 1.5e5 terms): 2e-4 of each block's largest entry."""
    from apnerf._lib import call
    from apnerf.radiance_fields import ngp

    g = torch.Generator().manual_seed(n)
    n_pad = ngp._padded_rows(n)
    X = torch.zeros((n_pad, ngp._X_WIDTH), dtype=torch.float16)
    G = torch.zeros((n_pad, ngp._G_WIDTH), dtype=torch.float16)
    X[:n] = torch.randn((n, ngp._X_WIDTH), generator=g).clamp_min(0).half()       # ReLU-like activations
    G[:n] = (torch.randn((n, ngp._G_WIDTH), generator=g) * 0.05).half()
    X, G = X.to(DEV), G.to(DEV)
    field = apnerf.NGPRadianceField([0, 0, 0, 1, 1, 1], layers=2, num_semantic_classes=C).to(DEV)
    n_sem_flat = sum(a * b for a, b in field._sem_dims_flat)
    out = [torch.zeros(field._n_base_w, device=DEV), torch.zeros(7168, device=DEV),
           torch.zeros(n_sem_flat, device=DEV) if C > 0 else None]
    ref = [torch.zeros_like(o) if o is not None else None for o in out]
    call("apnerf_field_wgrad", n, G, X, float(ngp.LOSS_SCALE), out[0], out[1], out[2],
         field._sem_dims_flat[2][0] if C > 0 else 32)
    ngp._wgrad_library(field, G, X, ref[0], ref[1], ref[2])
    torch.cuda.synchronize()
    dims = [field._base_dims, field._head_dims, field._sem_dims_flat if C > 0 else []]
    for o, r, dd in zip(out, ref, dims):
        if o is None:
            continue
        off = 0
        for a, b in dd:
            blk_o, blk_r = o[off:off + a * b], r[off:off + a * b]
            scale = float(blk_r.abs().max())
            assert scale > 0
            err = float((blk_o - blk_r).abs().max())
            assert err <= 2e-4 * scale + 1e-7, (n, C, (a, b), err, scale)
            off += a * b
