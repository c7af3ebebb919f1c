"""Per-kernel timing of the large-D vector field (tensor-core RFF term + RBF term) against the all-FP32 tiled kernel."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import gpode_oracle as O
from gaussian_process_odes_b200 import ops, _lib


def run(D, M, S, B):
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=5)
    gp = O.gp_params(p)
    omega = draws['eps_omega'] / gp['ell'].T.unsqueeze(1)
    nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.1, dtype=torch.float32)
    args = [a.cuda().contiguous() for a in (gp['Z'], gp['ell'], gp['var'], nu, omega, draws['phase_u'] * 2 * np.pi, draws['w'])]
    x = torch.randn(B, D, device="cuda")
    with torch.no_grad():
        field = ops.LargeField(*args)
        for _ in range(2):
            f = field(x)
        _lib.profile_start()
        for _ in range(3):
            f = field(x)
        prof = _lib.profile_stop()
        f_old = ops._large_d_call("gpode_vf_fwd_large", x, None, *args)
        _lib.profile_start()
        for _ in range(2):
            ops._large_d_call("gpode_vf_fwd_large", x, None, *args)
        prof.update(_lib.profile_stop())
    fv_rff, fv_rbf = D * S * (2 * D + 4), D * M * (3 * D + 4)
    out = dict(D=D, M=M, S=S, B=B, rel_diff=float((f - f_old).abs().max() / f_old.abs().max()))
    for k, (n, ms) in prof.items():
        out[k + "_ms"] = ms / n
    out["rff_tflops"] = B * fv_rff / (out["gpode_rff_fwd_large_ms"] * 1e-3) / 1e12
    rbf_key = "gpode_rbf_fwd_large_ms" if "gpode_rbf_fwd_large_ms" in out else "gpode_vf_fwd_large_add_rbf_ms"
    out["rbf_tflops"] = B * fv_rbf / (out[rbf_key] * 1e-3) / 1e12
    total = out["gpode_rff_fwd_large_ms"] + out[rbf_key]
    out["vf_total_ms"] = total
    out["vf_tflops"] = B * (fv_rff + fv_rbf) / (total * 1e-3) / 1e12
    out["speedup_vs_fp32_tiles"] = out["gpode_vf_fwd_large_ms"] / total
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)


if __name__ == "__main__":
    for D in (16, 32, 64):
        run(D, 100, 256, 100000)
    run(64, 100, 256, 10000)
    run(12, 50, 100, 100000)
