"""CPU: the host-side whitening route for 8 < D <= 64 (ops._whiten_large: float64 torch.linalg, differentiated by
autograd) against the oracle's restatement of DSVGP_Layer.build_cache (reference src/core/dsvgp.py:92-122). The public
``ops.whiten`` refuses CPU tensors (the product has no CPU path); only this pure-torch helper is exercised here, so the
large-D model classes' whitening is pinned without a GPU as well (tests/test_gpu_models.py covers it on the device)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pth in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, pth)


@pytest.mark.parametrize("D,M,S", [(9, 12, 16), (17, 20, 32), (64, 10, 8)])
def test_whiten_large_matches_the_oracle_build_cache(D, M, S):
    import gpode_oracle as O
    from gaussian_process_odes_b200 import ops
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=3, seed=D)
    cot = torch.tensor(np.random.default_rng(3).normal(size=(D, M)), dtype=torch.float32)

    def run(dtype, product):
        leaves = {k: p[k].detach().clone().to(dtype).requires_grad_(True)
                  for k in ("inducing_loc", "Um", "Us_sqrt_packed", "unconstrained_lengthscales",
                            "unconstrained_variance")}
        gp = O.gp_params(leaves)
        dr = {k: v.to(dtype) for k, v in draws.items() if torch.is_tensor(v) and v.is_floating_point()}
        cache = O.build_cache(gp['Z'], gp['Um'], gp['Us_sqrt'], gp['ell'], gp['var'], dr['w'], dr['eps_omega'],
                              dr['phase_u'], dr['eps_u'])
        if product:
            nu = ops._whiten_large(gp['Z'], gp['ell'], gp['var'], cache['u'], cache['rff_omega'],
                                   cache['rff_phase'], cache['rff_weights'], O.JITTER)      # (D, M) float32
        else:
            nu = cache['nu'].squeeze(2)
        (nu * cot.to(nu.dtype)).sum().backward()
        return nu.detach(), {k: v.grad.detach() for k, v in leaves.items()}

    nu64, g64 = run(torch.float64, False)
    nu32, g32 = run(torch.float32, False)
    nu, g = run(torch.float32, True)

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300))

    assert nu.dtype == torch.float32 and nu.shape == (D, M)
    # the helper computes in float64 from float32 parameters: at least as close to float64 as the float32 restatement
    assert rel(nu, nu64) <= max(1e-5, 1.5 * rel(nu32, nu64)), (rel(nu, nu64), rel(nu32, nu64))
    for k in g:
        assert rel(g[k], g64[k]) <= max(1e-4, 1.5 * rel(g32[k], g64[k])), (k, rel(g[k], g64[k]), rel(g32[k], g64[k]))
