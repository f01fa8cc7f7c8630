"""Prints the agreement of the fused training kernels' gradients with an fp32 torch restatement
(the numbers behind tests/test_training_gpu.py::test_field_gradients_match_fp32_reference) and a per-phase
timing of one training step.  GPU box:  python tools/train_grad_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import apnerf  # noqa: E402,F401
import test_training_gpu as T  # noqa: E402

real_cos = torch.nn.functional.cosine_similarity


def spy(a, b, dim=0):
    c = real_cos(a, b, dim=dim)
    print("  cosine %.6f  relative error %.4f  (|g| %.3e)" % (float(c), float((a - b).norm() / b.norm()), float(b.norm())))
    return c


torch.nn.functional.cosine_similarity = spy
T.test_field_gradients_match_fp32_reference(apnerf)
torch.nn.functional.cosine_similarity = real_cos

# timing of forward / backward of the field alone at a training-step-like size
f = T._field(apnerf).train()
n = 262144
g = torch.Generator().manual_seed(0)
lo, hi = torch.tensor(T.AABB[:3]), torch.tensor(T.AABB[3:])
pos = (lo + (hi - lo) * torch.rand((n, 3), generator=g)).to(T.DEV)
dirs = torch.randn((n, 3), generator=g)
dirs = (dirs / dirs.norm(dim=-1, keepdim=True)).to(T.DEV)
for it in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    rgb, dens, sem = f(pos, dirs)
    ev[1].record()
    (rgb.sum() + dens.sum() * 1e-3 + sem.sum() * 1e-2).backward()
    ev[2].record()
    torch.cuda.synchronize()
    print("n = %d: forward %.3f ms, backward %.3f ms" % (n, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
