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


def _grads(fn_out, leaves, cot):
    (fn_out * cot).sum().backward()
    return {k: v.grad.detach().cpu() for k, v in leaves.items()}


@pytest.mark.parametrize("D,M,S,B,ts", [
    (2, 16, 256, 1, [0.0, 0.3, 0.35, 1.0, 1.7]),
    (2, 16, 256, 125, [0.0, 0.29]),
    (5, 100, 256, 60, [0.0, 0.01, 0.02, 0.05]),
    (3, 24, 64, 300, [0.0, -0.2, -0.5]),
])
def test_dopri5_backward(D, M, S, B, ts):
    """Gradients of the adaptive solve against autograd through the restated torchdiffeq (float32 and float64)."""
    from gaussian_process_odes_b200 import ops
    gp, c, x, args = _problem(D, M, S, B, seed=D * 11 + B)
    ts = torch.tensor(ts, dtype=torch.float32)
    cot = torch.tensor(np.random.default_rng(3).normal(size=(len(ts), B, D)), dtype=torch.float32)
    for a in args[:4]:
        a.requires_grad_(True)
    xc = x.cuda().requires_grad_(True)
    xs, stats = ops.dopri5_integrate(xc, ts.cuda(), *args)
    (xs * cot.cuda()).sum().backward()
    got = dict(x=xc.grad.cpu(), Z=args[0].grad.cpu(), ell=args[1].grad.cpu(), var=args[2].grad.cpu(),
               nu=args[3].grad.cpu())
    res = {}
    for dtype in (torch.float32, torch.float64):
        leaves = dict(x=x.to(dtype).clone().requires_grad_(True), Z=gp['Z'].to(dtype).clone().requires_grad_(True),
                      ell=gp['ell'].to(dtype).clone().requires_grad_(True),
                      var=gp['var'].to(dtype).clone().requires_grad_(True),
                      nu=c['nu'].to(dtype).clone().requires_grad_(True))
        eps = (c['rff_omega'].double() * gp['ell'].double().T.unsqueeze(1)).to(dtype)
        cc = dict(rff_omega=eps / leaves['ell'].T.unsqueeze(1), rff_phase=c['rff_phase'].to(dtype),
                  rff_weights=c['rff_weights'].to(dtype), nu=leaves['nu'])
        out = O.odeint(lambda t, y: O.vf_forward(y, leaves['Z'], leaves['ell'], leaves['var'], cc), leaves['x'],
                       ts.to(dtype), method='dopri5', rtol=1e-6, atol=1e-6)
        res[dtype] = _grads(out, leaves, cot.to(dtype))
    for k in got:
        # both are exact gradients of a discrete solve at rtol=atol=1e-6; they may differ by one accept/reject decision
        assert_parity("dopri5 grad " + k, got[k].reshape(res[torch.float32][k].shape), res[torch.float32][k],
                      res[torch.float64][k], 2e-4)


def test_dopri5_training_through_flow():
    """solver='dopri5' (the reference default) trains: ELBO value and gradients against the golden reference run."""
    from util import build_product_model, injected_draws, load_golden, product_grads
    g = load_golden("vdp_shooting_dopri5")
    model = build_product_model("shooting", g['p'], g['ys'], 256, "dopri5")
    with injected_draws(g['draws'], mvn_order=("eps_x0", "eps_states")):
        ll, cst, e, k0 = model.build_lowerbound_terms(g['ys'].cuda(), g['ts'].cuda(), num_samples=5)
        loss = -(ll + cst + e - k0 - model.build_inducing_kl())
    loss.backward()
    assert_parity("loss", loss, g['ref']['loss'], g['f64']['loss'], 1e-4)
    for k, v in product_grads(model, "shooting").items():
        assert_parity("grad " + k, v.cpu(), g['ref']['grad_' + k], g['f64']['grad_' + k], 2e-4)
