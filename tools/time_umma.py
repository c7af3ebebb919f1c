"""Time the experimental tcgen05 vector field against the FFMA2 kernel (B rows, one evaluation)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import gpode_oracle as O
from util import oracle_cache, to_dev
from gaussian_process_odes_b200 import ops, _lib

def run(D, M, S, B):
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=3)
    gp, c = oracle_cache(p, draws)
    d = to_dev(dict(Z=gp['Z'], ell=gp['ell'], var=gp['var'], nu=c['nu'], omega=c['rff_omega'], phase=c['rff_phase'], w=c['rff_weights']))
    pc = ops.PackedCache(*[d[k].float().contiguous() for k in ("Z", "ell", "var", "nu", "omega", "phase", "w")])
    x = torch.randn(B, D, device="cuda"); f = torch.empty_like(x); f2 = torch.empty_like(x)
    lib = _lib.load(); st = _lib.stream_ptr()
    def t(name, out):
        for _ in range(3): _lib.call(name, _lib.ptr(pc.packed), D, M, S, _lib.ptr(x), _lib.ptr(out), B, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): _lib.call(name, _lib.ptr(pc.packed), D, M, S, _lib.ptr(x), _lib.ptr(out), B, st)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10
    a, b = t("gpode_vf_fwd", f), t("gpode_vf_fwd_umma", f2)
    err = float((f - f2).abs().max() / f.abs().max())
    print(json.dumps(dict(D=D, M=M, S=S, B=B, ffma2_ms=a, umma_ms=b, speedup=a / b, rel_diff=err)), flush=True)

if __name__ == "__main__":
    for D, M in ((2, 16), (3, 100), (4, 100), (5, 100), (6, 100), (7, 100)):
        run(D, M, 256, 1000000)
