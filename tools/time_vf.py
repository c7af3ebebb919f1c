#!/usr/bin/env python
"""Time the vector-field forward kernel at B=1e6 for a few state dimensions (tuning helper)."""
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

shapes = [(5, 100), (2, 16), (8, 100), (4, 100)] if len(sys.argv) < 2 else [(int(sys.argv[1]), int(sys.argv[2]))]
for D, M in shapes:
    S = 256
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=5)
    gp = O.gp_params(p)
    omega = draws["eps_omega"] / gp["ell"].T.unsqueeze(1)
    nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.1, dtype=torch.float32)
    args = [t.cuda().contiguous() for t in (gp["Z"], gp["ell"], gp["var"], nu, omega, draws["phase_u"] * 2 * np.pi,
                                            draws["w"])]
    B = 1000000
    x = torch.randn(B, D, device="cuda")
    for fmode in ("0", "1"):
        _lib.set_option("fwd_mma", int(fmode))
        tg = torch.tensor([0.0, 0.01], device="cuda")
        with torch.no_grad():
            for _ in range(3):
                ops.vector_field(x, *args)
                ops.rk4_integrate(x, tg, *args)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            for _ in range(5):
                ops.vector_field(x, *args)
            ev[1].record()
            for _ in range(5):
                ops.rk4_integrate(x, tg, *args)
            ev[2].record()
            torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        fv = D * (S * (2 * D + 4) + M * (3 * D + 4))
        print(json.dumps(dict(fwd_mma=fmode, D=D, ms=round(ms, 4), tflops=round(B * fv / (ms * 1e-3) / 1e12, 2),
                              rk4_step_ms=round(ev[1].elapsed_time(ev[2]) / 5, 4))), flush=True)
