#!/usr/bin/env python
"""Small driver for ncu: the RK4 forward / adjoint / param-grad kernels alone at a given shape.

    python tools/profile_kernels.py [--D 5 --M 100 --S 256 --B 1000000 --steps 1 --reps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpode_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=5)
ap.add_argument("--M", type=int, default=100)
ap.add_argument("--S", type=int, default=256)
ap.add_argument("--B", type=int, default=1000000)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
from gaussian_process_odes_b200 import ops  # noqa: E402

p, ys, ts, draws, _ = O.make_problem(D=a.D, M=a.M, S=a.S, N=1, T=4, seed=5)
gp = O.gp_params(p)
omega = draws['eps_omega'] / gp['ell'].T.unsqueeze(1)
nu = torch.tensor(np.random.default_rng(1).normal(size=(a.D, a.M)) * 0.1, dtype=torch.float32)
args = [t.cuda().contiguous() for t in (gp['Z'], gp['ell'], gp['var'], nu, omega, draws['phase_u'] * 2 * np.pi,
                                        draws['w'])]
for t in args[:4]:
    t.requires_grad_(True)
x = torch.randn(a.B, a.D, device="cuda", requires_grad=True)
tg = (torch.arange(a.steps + 1, dtype=torch.float32) * 0.01).cuda()
for _ in range(a.reps):
    xs = ops.rk4_integrate(x, tg, *args)
    xs.sum().backward()
torch.cuda.synchronize()
print("ok")
