#!/usr/bin/env python
"""Forward tensor-core kernel at B = 1e6: schedule variants of vf_eval_h (mma_parts 3 = fused RFF/RBF stream, the default;
7 = two parts; 11 = two parts, staggered across warps). Prints ms per evaluation / per RK4 step and the deviation from the staggered result."""
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
modes = [int(v) for v in os.environ.get("MODES", "3,7,11").split(",")]
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
    tg = torch.tensor([0.0, 0.01], device="cuda")
    ref = None
    for mode in modes:
        _lib.set_option("mma_parts", mode)
        with torch.no_grad():
            for _ in range(3):
                f = ops.vector_field(x, *args)
                xs = ops.rk4_integrate(x, tg, *args)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            for _ in range(10):
                ops.vector_field(x, *args)
            ev[1].record()
            for _ in range(10):
                ops.rk4_integrate(x, tg, *args)
            ev[2].record()
            torch.cuda.synchronize()
        if ref is None:
            ref = (f.clone(), xs.clone())
        dev = float((f - ref[0]).abs().max() / ref[0].abs().max())
        fv = D * (S * (2 * D + 4) + M * (3 * D + 4))
        ms = ev[0].elapsed_time(ev[1]) / 10
        print(json.dumps(dict(D=D, M=M, mma_parts=mode, vf_ms=round(ms, 4), vf_tflops=round(B * fv / (ms * 1e-3) / 1e12, 2),
                              rk4_step_ms=round(ev[1].elapsed_time(ev[2]) / 10, 4), max_rel_dev_vs_first=dev)), flush=True)
_lib.set_option("mma_parts", 3)
