#!/usr/bin/env python
"""One vector-field forward + backward at 8 < D <= 64 (for `ncu --metrics gpu__time_duration.sum`: per-kernel durations of
the large-D VJP -- tensor-core RFF kernel vs FP32 RBF kernel).   python tools/large_vjp_once.py D [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import gpode_oracle as O  # noqa: E402
from gaussian_process_odes_b200 import ops  # noqa: E402

D = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
M, S = 100, 256
p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=5)
gp = O.gp_params(p)
omega = draws['eps_omega'] / gp['ell'].T.unsqueeze(1)
nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.1, dtype=torch.float32)
args = [a.cuda().contiguous() for a in (gp['Z'], gp['ell'], gp['var'], nu, omega, draws['phase_u'] * 2 * np.pi, draws['w'])]
for a in args[:4]:
    a.requires_grad_(True)
x = torch.randn(B, D, device="cuda", requires_grad=True)
cot = torch.randn(B, D, device="cuda")
for _ in range(3):
    ops.vector_field(x, *args).backward(cot)
torch.cuda.synchronize()
print("done", D, B)
