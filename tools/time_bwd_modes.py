#!/usr/bin/env python
"""RK4 adjoint (one step, B = 1e6) with the tensor-core kernel: schedule variants of vf_vjp_h (mma_parts 3 = two parts,
half of the warps in the opposite order; 19 = fused RFF/RBF stream, only in a build with
-DGPODE_VJP_FUSED_EXPERIMENT). Prints ms per backward pass (adjoint + gradient
contraction + finalize) and the deviation of every gradient from the first variant."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpode_oracle as O  # noqa: E402
from gaussian_process_odes_b200 import ops, _lib  # noqa: E402

shapes = [(5, 100), (4, 100)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
modes = [int(v) for v in os.environ.get("MODES", "3,19").split(",")]
B = int(os.environ.get("TB_ROWS", 1000000))
for D, M in shapes:
    S = 256
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=5)
    gp = O.gp_params(p)
    omega = draws["eps_omega"] / gp["ell"].T.unsqueeze(1)
    nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.1, dtype=torch.float32)
    base = [t.cuda().contiguous() for t in (gp["Z"], gp["ell"], gp["var"], nu, omega, draws["phase_u"] * 2 * np.pi,
                                            draws["w"])]
    x = torch.randn(B, D, device="cuda")
    cot = torch.randn(2, B, D, device="cuda") / B
    tg = torch.tensor([0.0, 0.01], device="cuda")
    first = None
    for mode in modes:
        _lib.set_option("mma_parts", mode)
        args = [a.detach().clone().requires_grad_(i < 4) for i, a in enumerate(base)]
        xc = x.clone().requires_grad_(True)
        xs = ops.rk4_integrate(xc, tg, *args)
        for _ in range(3):
            xs.backward(cot, retain_graph=True)
        for a in [xc] + args[:4]:
            a.grad = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            xs.backward(cot, retain_graph=True)
        e1.record()
        torch.cuda.synchronize()
        g = [(a.grad / n).clone() for a in [xc] + args[:4]]
        if first is None:
            first = g
        rel = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(g, first)]
        print(json.dumps(dict(D=D, M=M, B=B, mma_parts=mode, bwd_ms=round(e0.elapsed_time(e1) / n, 3),
                              rel_x_Z_ell_var_nu_vs_first=["%.1e" % r for r in rel])), flush=True)
_lib.set_option("mma_parts", 3)
