"""-m gpu: the experimental tcgen05 (tensor-core) vector field against the FFMA2 kernel and the oracle."""
import numpy as np
import pytest
import torch

import gpode_oracle as O
from util import TOL_VF, assert_parity, oracle_cache, relerr, to_dev

pytestmark = pytest.mark.gpu


def _problem(D, M, S, B, seed, nu_scale=0.3):
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=seed)
    gp, c = oracle_cache(p, draws)
    c['nu'] = torch.tensor(np.random.default_rng(seed).normal(size=(D, M, 1)) * nu_scale, dtype=torch.float32)
    x = torch.tensor(np.random.default_rng(seed + 1).normal(size=(B, D)) * 1.5, dtype=torch.float32)
    d = to_dev(dict(Z=gp['Z'], ell=gp['ell'], var=gp['var'], nu=c['nu'], omega=c['rff_omega'], phase=c['rff_phase'],
                    w=c['rff_weights']))
    args = [d[k].float().contiguous() for k in ("Z", "ell", "var", "nu", "omega", "phase", "w")]
    return gp, c, x, args


@pytest.mark.parametrize("D,M,S,B", [(2, 16, 256, 1), (2, 16, 256, 1000), (5, 100, 256, 4097), (3, 24, 40, 333),
                                      (7, 30, 300, 129), (5, 100, 256, 200000), (4, 50, 512, 5000)])
def test_vf_umma_matches_ffma2_and_oracle(D, M, S, B):
    from gaussian_process_odes_b200 import ops
    gp, c, x, args = _problem(D, M, S, B, seed=D * 7 + S)
    with torch.no_grad():
        f_tc = ops.vector_field_umma(x.cuda(), *args)
        f_fma = ops.vector_field(x.cuda(), *args)
    torch.cuda.synchronize()
    assert relerr(f_tc, f_fma) <= 3e-6, "tensor-core theta differs from the FFMA2 kernel"
    n = min(B, 3000)
    ref = O.vf_forward(x[:n], gp['Z'], gp['ell'], gp['var'], c)
    gp64 = {k: v.double() for k, v in gp.items()}
    c64 = {k: v.double() for k, v in c.items()}
    ref64 = O.vf_forward(x[:n].double(), gp64['Z'], gp64['ell'], gp64['var'], c64)
    assert_parity("vf_umma", f_tc[:n].cpu(), ref, ref64, TOL_VF)


def test_vf_backward_tcgen05_experiment_matches_ffma2():
    """csrc/vjp_umma.cu (measured experiment, not dispatched): the VJP with theta as a kind::f16 tcgen05 GEMM, the sine
    written back to tensor memory as the TS-form A operand of the second GEMM. Same grad_x and parameter gradients as
    the FFMA2 adjoint (gpode_vf_bwd with bwd_mma = 0)."""
    import ctypes
    import numpy as np
    import gpode_oracle as O
    from gaussian_process_odes_b200 import _lib, ops
    from gaussian_process_odes_b200._lib import ptr, stream_ptr
    D, M, S, B = 5, 33, 100, 40000
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=3)
    gp = O.gp_params(p)
    omega = draws["eps_omega"] / gp["ell"].T.unsqueeze(1)
    nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.3, dtype=torch.float32)
    args = [t.cuda().contiguous() for t in (gp["Z"], gp["ell"], gp["var"], nu, omega, draws["phase_u"] * 2 * np.pi,
                                            draws["w"])]
    pc = ops.PackedCache(*args)
    lib = _lib.load()
    ub = torch.empty(lib.gpode_packed_ubwd_floats(D, S), dtype=torch.float32, device="cuda")
    _lib.call("gpode_pack_cache_ubwd", ctypes.byref(pc.struct), ptr(ub), stream_ptr())
    x = torch.randn(B, D, device="cuda") * 1.5
    gf = torch.randn(B, D, device="cuda")
    f = torch.empty_like(x)
    _lib.call("gpode_vf_fwd", ptr(pc.packed), D, M, S, ptr(x), ptr(f), B, stream_ptr())
    res = {}
    for kind in ("ffma2", "umma"):
        gx, acc = torch.empty_like(x), pc.new_acc()
        if kind == "umma":
            _lib.call("gpode_vf_bwd_umma", ptr(pc.packed), ptr(ub), D, M, S, ptr(x), ptr(f), ptr(gf), ptr(gx), ptr(acc),
                      B, stream_ptr())
        else:
            _lib.set_option("bwd_mma", 0)
            try:
                _lib.call("gpode_vf_bwd", ptr(pc.packed), D, M, S, ptr(x), ptr(f), ptr(gf), ptr(gx), ptr(acc), B,
                          stream_ptr())
            finally:
                _lib.set_option("bwd_mma", 1)
        res[kind] = (gx, pc.finalize(acc))
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(res["umma"][0], res["ffma2"][0]) <= 1e-5
    for a, b in zip(res["umma"][1][1:3], res["ffma2"][1][1:3]):   # lengthscale and variance gradients
        assert rel(a, b) <= 1e-5
