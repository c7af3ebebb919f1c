#!/usr/bin/env python
"""Time the RK4 adjoint (one step, B = 1e6) with the FFMA2 kernel (GPODE_BWD_MMA=0) and the tensor-core kernel
(default), and compare their gradients (tuning helper; prints one JSON line per state dimension)."""
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

shapes = [(5, 100), (4, 100), (7, 100), (3, 24), (6, 100)] if len(sys.argv) < 2 else [(int(sys.argv[1]), int(sys.argv[2]))]
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
    out = {}
    for mode in ("0", "1"):
        _lib.set_option("bwd_mma", int(mode))
        args = [a.detach().clone().requires_grad_(i < 4) for i, a in enumerate(base)]
        xc = x.clone().requires_grad_(True)
        xs = ops.rk4_integrate(xc, tg, *args)
        for _ in range(3):
            xs.backward(cot, retain_graph=True)
        for a in [xc] + args[:4]:
            a.grad = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            xs.backward(cot, retain_graph=True)
        e1.record()
        torch.cuda.synchronize()
        out[mode] = dict(ms=e0.elapsed_time(e1) / n, g=[(a.grad / n).clone() for a in [xc] + args[:4]])
    rel = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(out["1"]["g"], out["0"]["g"])]
    print(json.dumps(dict(D=D, M=M, B=B, bwd_ms_ffma2=round(out["0"]["ms"], 3), bwd_ms_mma=round(out["1"]["ms"], 3),
                          rel_x_Z_ell_var_nu=["%.2e" % r for r in rel])), flush=True)
