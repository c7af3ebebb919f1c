"""-m gpu: BASELINE.json's FULL size (10^6 MoCap-shaped rows: D=5, M=100, S=256) through size-independent properties:
rows are independent, so (a) any sample of rows must match the oracle run on just those rows, (b) permuting the batch
permutes the result bit for bit, (c) integrating [t0,t1,t2] equals integrating [t0,t1] then [t1,t2] bit for bit,
(d) the adjoint's shared-parameter gradients of the whole batch equal the sum over two half batches ("checksum of
checksums"), and the per-row gradients are those of the halves."""
import numpy as np
import pytest
import torch

import gpode_oracle as O
from util import TOL_TRAJ, TOL_VF, assert_parity, oracle_cache, relerr, to_dev

pytestmark = pytest.mark.gpu

D, M, S, B = 5, 100, 256, 1_000_000


@pytest.fixture(scope="module")
def problem():
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=77, ell0=1.25)
    gp, c = oracle_cache(p, draws)
    x = torch.tensor(np.random.default_rng(78).normal(size=(B, D)), dtype=torch.float32)
    d = to_dev(dict(Z=gp['Z'], ell=gp['ell'], var=gp['var'], nu=c['nu'], omega=c['rff_omega'], phase=c['rff_phase'],
                    w=c['rff_weights']))
    args = [d[k].float().contiguous() for k in ("Z", "ell", "var", "nu", "omega", "phase", "w")]
    return gp, c, x, args


def _oracle_rows(gp, c, x, idx, ts=None):
    xs = x[idx]
    f32 = lambda t, y: O.vf_forward(y, gp['Z'], gp['ell'], gp['var'], c)
    gp64 = {k: v.double() for k, v in gp.items()}
    c64 = {k: v.double() for k, v in c.items()}
    f64 = lambda t, y: O.vf_forward(y, gp64['Z'], gp64['ell'], gp64['var'], c64)
    if ts is None:
        return f32(None, xs), f64(None, xs.double())
    return O.odeint(f32, xs, ts, method='rk4'), O.odeint(f64, xs.double(), ts.double(), method='rk4')


def test_full_size_vector_field_sampled_rows_and_permutation(problem):
    from gaussian_process_odes_b200 import ops
    gp, c, x, args = problem
    xg = x.cuda()
    with torch.no_grad():
        f = ops.vector_field(xg, *args)
        perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        fp = ops.vector_field(xg[perm].contiguous(), *args)
    assert torch.equal(fp, f[perm]), "a row's value depends on its position in the batch"
    idx = torch.tensor(np.random.default_rng(5).choice(B, 2000, replace=False))
    ref32, ref64 = _oracle_rows(gp, c, x, idx)
    assert_parity("full-size vf sample", f[idx.cuda()].cpu(), ref32, ref64, TOL_VF)


def test_full_size_rk4_sampled_rows_and_composition(problem):
    from gaussian_process_odes_b200 import ops
    gp, c, x, args = problem
    xg = x.cuda()
    ts = torch.tensor([0.0, 0.01, 0.03], dtype=torch.float32)
    with torch.no_grad():
        xs = ops.rk4_integrate(xg, ts.cuda(), *args)                      # (3,B,D)
        a = ops.rk4_integrate(xg, ts[:2].cuda(), *args)
        b = ops.rk4_integrate(a[1].contiguous(), ts[1:].cuda(), *args)
    assert torch.equal(xs[0], xg) and torch.equal(xs[1], a[1]) and torch.equal(xs[2], b[1])
    idx = torch.tensor(np.random.default_rng(6).choice(B, 1000, replace=False))
    ref32, ref64 = _oracle_rows(gp, c, x, idx, ts)
    assert_parity("full-size rk4 sample", xs[:, idx.cuda()].cpu(), ref32, ref64, TOL_TRAJ)


def test_full_size_adjoint_is_additive_over_half_batches(problem):
    from gaussian_process_odes_b200 import ops
    gp, c, x, args = problem
    ts = torch.tensor([0.0, 0.01], dtype=torch.float32).cuda()
    cot = torch.tensor(np.random.default_rng(9).normal(size=(B, D)), dtype=torch.float32).cuda() / B

    def run(lo, hi):
        leaves = [a.detach().clone().requires_grad_(True) for a in args[:4]]
        x0 = x[lo:hi].cuda().requires_grad_(True)
        xs = ops.rk4_integrate(x0, ts, *leaves, *args[4:])
        (xs[-1] * cot[lo:hi]).sum().backward()
        return x0.grad, [l.grad for l in leaves]

    gx, gfull = run(0, B)
    gx1, g1 = run(0, B // 2)
    gx2, g2 = run(B // 2, B)
    # per-row gradients do not depend on the batch they were computed in
    assert torch.equal(gx[:B // 2], gx1) and torch.equal(gx[B // 2:], gx2)
    for name, a, b1, b2 in zip(("Z", "ell", "var", "nu"), gfull, g1, g2):
        assert relerr(a, b1 + b2) <= 1e-4, name
    # and a sample of rows against the oracle's autograd
    idx = torch.tensor(np.random.default_rng(10).choice(B, 500, replace=False))
    xo = x[idx].clone().requires_grad_(True)
    ref = O.odeint(lambda t, y: O.vf_forward(y, gp['Z'], gp['ell'], gp['var'], c), xo, ts.cpu(), method='rk4')
    (ref[-1] * cot[idx.cuda()].cpu()).sum().backward()
    assert relerr(gx[idx.cuda()].cpu(), xo.grad) <= 1e-4
