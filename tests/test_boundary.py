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
