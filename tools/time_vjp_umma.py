"""The tcgen05 / TMEM adjoint experiment (csrc/vjp_umma.cu) against the mma.sync adjoint (csrc/vjp_mma.cuh) and the FFMA2
adjoint on one VJP of B rows: parity of grad_x and of the four parameter gradients, and CUDA-event timings."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import gpode_oracle as O
from gaussian_process_odes_b200 import ops, _lib
from gaussian_process_odes_b200._lib import ptr, stream_ptr

D, M, S = int(os.environ.get("TD", 5)), 100, 256
B = int(os.environ.get("TB_ROWS", 1000000))
p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=5)
gp = O.gp_params(p)
omega = draws["eps_omega"] / gp["ell"].T.unsqueeze(1)
nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.1, dtype=torch.float32)
args = [t.cuda().contiguous() for t in (gp["Z"], gp["ell"], gp["var"], nu, omega, draws["phase_u"] * 2 * np.pi, draws["w"])]
pc = ops.PackedCache(*args)
lib = _lib.load()
ub = torch.empty(lib.gpode_packed_ubwd_floats(D, S), dtype=torch.float32, device="cuda")
_lib.call("gpode_pack_cache_ubwd", ctypes.byref(pc.struct), ptr(ub), stream_ptr())
x = torch.randn(B, D, device="cuda")
gf = torch.randn(B, D, device="cuda") / B
f = torch.empty_like(x)
_lib.call("gpode_vf_fwd", ptr(pc.packed), D, M, S, ptr(x), ptr(f), B, stream_ptr())


def run(kind, reps=5):
    gx = torch.empty_like(x)
    def once():
        acc = pc.new_acc()
        if kind == "umma":
            _lib.call("gpode_vf_bwd_umma", ptr(pc.packed), ptr(ub), D, M, S, ptr(x), ptr(f), ptr(gf), ptr(gx), ptr(acc), B,
                      stream_ptr())
        else:
            _lib.set_option("bwd_mma", 1 if kind == "mma" else 0)
            _lib.call("gpode_vf_bwd", ptr(pc.packed), D, M, S, ptr(x), ptr(f), ptr(gf), ptr(gx), ptr(acc), B, stream_ptr())
        return acc
    acc = once()
    torch.cuda.synchronize()
    tt = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2 = torch.empty_like(acc); a2[:4].zero_()
        e0.record()
        if kind == "umma":
            _lib.call("gpode_vf_bwd_umma", ptr(pc.packed), ptr(ub), D, M, S, ptr(x), ptr(f), ptr(gf), ptr(gx), ptr(a2), B,
                      stream_ptr())
        else:
            _lib.call("gpode_vf_bwd", ptr(pc.packed), D, M, S, ptr(x), ptr(f), ptr(gf), ptr(gx), ptr(a2), B, stream_ptr())
        e1.record(); torch.cuda.synchronize()
        tt.append(e0.elapsed_time(e1))
    g = pc.finalize(acc)
    return float(np.median(tt)), gx.clone(), [t.clone() for t in g]


rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
out = {"D": D, "B": B}
ms0, gx0, g0 = run("ffma2")
ms1, gx1, g1 = run("mma")
out.update(ffma2_ms=ms0, mma_sync_ms=ms1, mma_vs_ffma2_gx=rel(gx1, gx0))
try:
    ms2, gx2, g2 = run("umma")
    out.update(umma_ms=ms2, umma_vs_ffma2_gx=rel(gx2, gx0), umma_vs_ffma2_ell=rel(g2[1], g0[1]),
               umma_vs_ffma2_var=rel(g2[2], g0[2]), nan=bool(torch.isnan(gx2).any()))
except Exception as e:
    out["umma_error"] = repr(e)[:300]
_lib.set_option("bwd_mma", 1)
print(json.dumps(out))
