"""Diagnostics for the fused field kernel on a GPU box: prints error statistics of every output
against the CPU oracle (used while bringing up the tcgen05 path)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apnerf  # noqa: E402
from apnerf import synthetic  # noqa: E402
from oracle import oracle as O  # noqa: E402

AABB = [-6.4, -0.2, -6.4, 6.4, 12.6, 6.4]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
f = apnerf.NGPRadianceField(AABB, layers=2, num_semantic_classes=29)
synthetic.init_trained_like(f, seed=2)
f = f.to("cuda:0").eval()
fp = O.FieldParams(f.mlp_base.params.detach().cpu().numpy(), f.mlp_head.params.detach().cpu().numpy(),
                   f.mlp_sem.params.detach().cpu().numpy(), num_semantic_classes=29)
g = torch.Generator().manual_seed(0)
lo, hi = torch.tensor(AABB[:3]), torch.tensor(AABB[3:])
pos = lo + (hi - lo) * torch.rand((n, 3), generator=g)
dirs = torch.randn((n, 3), generator=g)
dirs = dirs / dirs.norm(dim=-1, keepdim=True)
with torch.no_grad():
    d, feat = f.query_density(pos.cuda(), return_feat=True)
    torch.cuda.synchronize()
    print("density-only ok", d.shape, float(d.mean()))
    rgb, dens, sem = f(pos.cuda(), dirs.cuda())
    torch.cuda.synchronize()
orgb, odens, osem = O.field_forward(pos.numpy(), dirs.numpy(), np.asarray(AABB, np.float32), fp)
x = ((pos - lo) / (hi - lo)).numpy()
enc = O.hashgrid_encode(x, fp.table, fp.meta)
base = O.mlp_forward(enc, fp.base_w).astype(np.float32)
print("feat  max|err|", np.abs(feat.float().cpu().numpy() - base[:, 1:16]).max(), "scale", np.abs(base).max())
print("dens  rel err  median/max", np.median(np.abs(dens.cpu().numpy() - odens) / np.maximum(odens, 1e-6)),
      (np.abs(dens.cpu().numpy() - odens) / np.maximum(odens, 1e-6)).max())
print("rgb   max|err|", np.abs(rgb.cpu().numpy() - orgb).max())
print("sem   max|err|", np.abs(sem.cpu().numpy() - osem).max(), "scale", np.abs(osem).max())
print("sample rows:\n", rgb[:3].cpu().numpy(), "\n", orgb[:3])
