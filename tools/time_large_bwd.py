"""Timing of the differentiable large-D path (8 < D <= 64): vector-field VJP kernel, one RK4 step forward + adjoint, and
the device-side dopri5, on the shapes of BASELINE.json configs[4] (nu supplied directly, as the sweep specifies).

    python tools/time_large_bwd.py [--out profiles/r02_large_d.json]
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import gpode_oracle as O
from gaussian_process_odes_b200 import ops, _lib


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def run(D, M, S, B):
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=5)
    gp = O.gp_params(p)
    omega = draws['eps_omega'] / gp['ell'].T.unsqueeze(1)
    nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.1, dtype=torch.float32)
    args = [a.cuda().contiguous() for a in (gp['Z'], gp['ell'], gp['var'], nu, omega, draws['phase_u'] * 2 * np.pi, draws['w'])]
    for a in args[:4]:
        a.requires_grad_(True)
    x = torch.randn(B, D, device="cuda", requires_grad=True)
    fv = D * (S * (2 * D + 4) + M * (3 * D + 4))
    out = dict(D=D, M=M, S=S, B=B, F_vf=fv)
    cot = torch.randn(B, D, device="cuda")
    f = ops.vector_field(x, *args)
    out["vf_fwd_ms"] = ev_time(lambda: ops.vector_field(x, *args))
    _lib.profile_start()
    for _ in range(3):
        f = ops.vector_field(x, *args)
        f.backward(cot)
    prof = _lib.profile_stop(raw=True)
    out["vjp_kernel_ms"] = float(np.median(prof["gpode_vf_bwd_large"]))
    out["vjp_tflops_algorithmic"] = B * 2 * fv / (out["vjp_kernel_ms"] * 1e-3) / 1e12
    tg = (torch.arange(2, dtype=torch.float32) * 0.01).cuda()
    cot2 = torch.randn(2, B, D, device="cuda")

    def step():
        xs = ops.rk4_integrate(x, tg, *args)
        xs.backward(cot2)
    out["rk4_step_fwd_bwd_ms"] = ev_time(step)
    out["rk4_step_evals_per_s"] = 4 * B / (out["rk4_step_fwd_bwd_ms"] * 1e-3)
    with torch.no_grad():
        t5 = torch.tensor([0.0, 0.16, 0.32], dtype=torch.float32).cuda()
        xs, stats = ops.dopri5_integrate(x.detach(), t5, *[a.detach() for a in args])
        out["dopri5_ms"] = ev_time(lambda: ops.dopri5_integrate(x.detach(), t5, *[a.detach() for a in args]))
        out["dopri5_stats_nfe_acc_rej_status"] = [int(v) for v in stats.cpu()]
        out["dopri5_evals_per_s"] = out["dopri5_stats_nfe_acc_rej_status"][0] * B / (out["dopri5_ms"] * 1e-3)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = []
    shapes = [(16, 100, 256, 100000), (32, 100, 256, 100000), (64, 100, 256, 100000), (16, 100, 256, 1000000),
              (32, 100, 256, 1000000), (64, 100, 256, 1000000)]
    if os.environ.get("TL_SMALL"):   # quick A/B of the kernel-selection options: 1e5 rows only
        shapes = shapes[:3]
    for D, M, S, B in shapes:
        r = run(D, M, S, B)
        print(json.dumps(r), flush=True)
        res.append(r)
        torch.cuda.empty_cache()
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
