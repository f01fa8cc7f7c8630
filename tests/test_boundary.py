"""CPU: the C-ABI library loads and exports every symbol include/apnerf.h declares; host-side
packing logic.  No compute calls (no GPU here)."""
import ctypes
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(apnerf):
    protos = apnerf._lib.parse_header()
    assert len(protos) >= 15
    assert os.path.exists(apnerf._lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    dll = ctypes.CDLL(apnerf._lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), f"libapnerf.so does not export {name}"
    assert apnerf._lib.LIB.raw("apnerf_abi_version")() == 1
    assert apnerf._lib.LIB.raw("apnerf_field_weight_bytes")() == 81920


def test_no_cpu_fallback(apnerf):
    """Product ops refuse CPU tensors instead of silently computing elsewhere."""
    from apnerf import nerfacc

    o = torch.zeros(4, 3)
    d = torch.ones(4, 3)
    aabbs = torch.tensor([[0.0, 0, 0, 1, 1, 1]])
    with pytest.raises(RuntimeError):
        nerfacc.ray_aabb_intersect(o, d, aabbs)
    with pytest.raises(NotImplementedError):
        nerfacc.pack_info(torch.tensor([0, 1, 1]), 2)  # reference behaviour, pack.py:48


def test_level_table_matches_oracle(apnerf, oracle):
    from apnerf.radiance_fields.ngp import hashgrid_levels

    meta, total = hashgrid_levels(16, 16, 4096, 19)
    ometa, ototal = oracle.hashgrid_meta()
    assert total == ototal == 6299960
    assert (meta[:, :4] == ometa).all()
    assert meta[:, 4].tolist() == [0] * 5 + [1] * 11


def test_field_state_dict_layout(apnerf):
    f = apnerf.NGPRadianceField([-1, -1, -1, 1, 1, 1], layers=2, num_semantic_classes=29)
    sd = f.state_dict()
    assert list(sd) == ["aabb", "direction_encoding.params", "mlp_base.params", "mlp_head.params", "mlp_sem.params"]
    assert sd["mlp_base.params"].numel() == 26624 + 25199840
    assert sd["mlp_head.params"].numel() == 7168 and sd["mlp_sem.params"].numel() == 7168
    assert sd["direction_encoding.params"].numel() == 0
    with pytest.raises(NotImplementedError):
        apnerf.NGPRadianceField([-1, -1, -1, 1, 1, 1], layers=4)


def test_umma_weight_packing(apnerf):
    from apnerf.radiance_fields.ngp import _umma_pack

    w = torch.arange(16 * 32, dtype=torch.float32).reshape(16, 32).to(torch.float16)
    p = _umma_pack(w)
    for n, k in [(0, 0), (3, 5), (15, 31), (7, 8), (8, 17)]:
        assert p[(k // 8) * (16 * 8) + n * 8 + k % 8] == w[n, k]


def test_training_matrix_layouts_are_consistent():
    """Host-side bookkeeping of the training path: the column blocks of the two wide fp16 matrices tile them exactly,
    every weight-gradient block pairs a gradient with the activation that produced it (shapes of the flat parameter
    vectors), and row padding is a whole number of GEMM chunks."""
    import apnerf
    from apnerf.radiance_fields import ngp

    for layout, width in ((ngp._X_COLS, ngp._X_WIDTH), (ngp._G_COLS, ngp._G_WIDTH)):
        spans = sorted(layout.values())
        assert spans[0][0] == 0 and spans[-1][1] == width
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % 8 == 0 for lo, _ in spans)  # 16-byte aligned column blocks
    f = apnerf.NGPRadianceField([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], layers=2, num_semantic_classes=29)
    dims = f._base_dims + f._head_dims + f._sem_dims
    assert len(dims) == len(ngp._WGRAD_BLOCKS) == 9
    for (n_out, n_in), (g, x) in zip(dims, ngp._WGRAD_BLOCKS):
        assert ngp._G_COLS[g][1] - ngp._G_COLS[g][0] == n_out, (g, n_out)
        assert ngp._X_COLS[x][1] - ngp._X_COLS[x][0] == n_in, (x, n_in)
    for n in (0, 1, ngp.WGRAD_CHUNK - 1, ngp.WGRAD_CHUNK, ngp.WGRAD_CHUNK + 1, 5 * ngp.WGRAD_CHUNK):
        p = ngp._padded_rows(n)
        assert p >= max(n, 1) and p % ngp.WGRAD_CHUNK == 0 and p - n < ngp.WGRAD_CHUNK + (n == 0)


def test_view_selection_matches_reference_indexing_for_short_trajectories():
    """pipeline.py:687-697 indexes trajectory[unc_idx] with numpy semantics: 40 indices with repeats, and negative
    ones (trajectories shorter than 20 poses) wrap around.  The scorer must pick the same poses."""
    import numpy as np

    from apnerf.scoring import uncertainty_view_indices

    for n in (10, 19, 20, 21, 24, 40, 41, 100):  # below 10 poses the reference itself raises IndexError
        traj = np.arange(n * 7, dtype=np.float64).reshape(n, 7)
        a = np.linspace(0, n - 20, 20)
        b = np.linspace(n - 20, n - 1, 20)
        ref = traj[np.hstack((a, b)).astype(int)]
        got = traj[uncertainty_view_indices(n)]
        assert got.shape == (40, 7) and np.array_equal(got, ref)


def _cuobjdump(*args):
    import shutil
    import subprocess

    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    return subprocess.run([exe, *args], capture_output=True, text=True, check=True).stdout


def test_built_kernels_match_the_documented_design(apnerf):
    """The shipped libapnerf.so is an sm_100a build whose hot kernels have the register budgets DESIGN.md argues
    from, and whose field / backward / weight-gradient kernels really are tcgen05 + TMEM code (UTCHMMA = tcgen05.mma,
    LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit): a recompiled mma.sync kernel would not contain them."""
    import re

    usage = _cuobjdump("-res-usage", apnerf._lib.LIB_PATH)
    assert "sm_100a" in usage
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", usage):
        regs[m.group(1)] = int(m.group(2))
    pick = lambda frag: [v for k, v in regs.items() if frag in k]
    assert pick("field_forward_kernel") and max(pick("field_forward_kernel")) <= 72  # 832 threads x 72 <= 65 536
    assert max(pick("render_composite_kernel")) <= 64  # 8 CTAs of 128 threads per SM
    assert max(v for k, v in regs.items() if "render_march_kernel" in k) <= 64  # 4 CTAs of 256 threads per SM
    assert max(pick("field_wgrad_kernel")) <= 64 and max(pick("field_backward_kernel")) <= 80
    sass = _cuobjdump("-sass", apnerf._lib.LIB_PATH)
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTCBAR"):
        assert sass.count(mnemonic) > 0, f"no {mnemonic} in the shipped SASS: the tcgen05 path is missing"
