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
