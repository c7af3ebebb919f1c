"""-m gpu: the cooperative adaptive integrator against the restated torchdiffeq 0.2.0 dopri5."""
import numpy as np
import pytest
import torch

import gpode_oracle as O
from util import TOL_TRAJ, assert_parity, oracle_cache, relerr, to_dev

pytestmark = pytest.mark.gpu


def _problem(D, M, S, B, seed):
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=seed)
    gp, c = oracle_cache(p, draws)
    c['nu'] = torch.tensor(np.random.default_rng(seed).normal(size=(D, M, 1)) * 0.3, dtype=torch.float32)
    x = torch.tensor(np.random.default_rng(seed + 1).normal(size=(B, D)), dtype=torch.float32)
    d = to_dev(dict(Z=gp['Z'], ell=gp['ell'], var=gp['var'], nu=c['nu'], omega=c['rff_omega'], phase=c['rff_phase'],
                    w=c['rff_weights']))
    args = [d[k].float().contiguous() for k in ("Z", "ell", "var", "nu", "omega", "phase", "w")]
    return gp, c, x, args


@pytest.mark.parametrize("D,M,S,B", [(2, 16, 256, 1), (2, 16, 256, 125), (5, 100, 256, 700), (3, 24, 64, 40000),
                                      (8, 30, 64, 33)])
@pytest.mark.parametrize("tol", [1e-6, 1e-4])
def test_dopri5_forward(D, M, S, B, tol):
    from gaussian_process_odes_b200 import ops
    gp, c, x, args = _problem(D, M, S, B, seed=D + B)
    ts = torch.tensor([0.0, 0.13, 0.5, 0.51, 1.2], dtype=torch.float32)
    xs, stats = ops.dopri5_integrate(x.cuda(), ts.cuda(), *args, rtol=tol, atol=tol)
    st = {}
    ref32 = O.odeint(lambda t, y: O.vf_forward(y, gp['Z'], gp['ell'], gp['var'], c), x, ts, method='dopri5', rtol=tol,
                     atol=tol, stats=st)
    gp64 = {k: v.double() for k, v in gp.items()}
    c64 = {k: v.double() for k, v in c.items()}
    ref64 = O.odeint(lambda t, y: O.vf_forward(y, gp64['Z'], gp64['ell'], gp64['var'], c64), x.double(), ts.double(),
                     method='dopri5', rtol=tol * 1e-3, atol=tol * 1e-3)
    nfe, acc, rej, status = [int(v) for v in stats.cpu()]
    assert status == 0
    assert torch.equal(xs[0].cpu(), x)
    # both solve the ODE to `tol`; they agree with each other to trajectory tolerance, arbitrated by a tight solve
    assert_parity("dopri5", xs.cpu(), ref32, ref64, max(TOL_TRAJ, 20 * tol))
    assert abs(acc - st['accepted']) <= 1 and abs(rej - st['rejected']) <= 1
    assert nfe == 2 + 6 * (acc + rej)


def test_dopri5_decreasing_grid_and_single_point():
    from gaussian_process_odes_b200 import ops
    gp, c, x, args = _problem(2, 16, 64, 9, seed=5)
    ts = torch.tensor([0.0, -0.2, -0.7], dtype=torch.float32)
    xs, stats = ops.dopri5_integrate(x.cuda(), ts.cuda(), *args)
    ref = O.odeint(lambda t, y: O.vf_forward(y, gp['Z'], gp['ell'], gp['var'], c), x, ts, method='dopri5', rtol=1e-6,
                   atol=1e-6)
    assert relerr(xs.cpu(), ref) <= TOL_TRAJ
    xs1, _ = ops.dopri5_integrate(x.cuda(), ts[:1].cuda(), *args)
    assert torch.equal(xs1[0].cpu(), x)


def test_dopri5_refuses_to_train_silently():
    from gaussian_process_odes_b200 import ops, _lib
    gp, c, x, args = _problem(2, 16, 64, 9, seed=5)
    args[0].requires_grad_(True)
    with pytest.raises(_lib.GpodeError):
        ops.dopri5_integrate(x.cuda(), torch.tensor([0.0, 0.1]).cuda(), *args)
