"""CPU: the oracle port against the committed golden fixtures (outputs of the UNMODIFIED reference, produced by
oracle/pin_against_reference.py in the build container)."""
import ast

import pytest
import torch

import gpode_oracle as O
from util import TOL_GRAD, TOL_TRAJ, assert_parity, load_golden, relerr

CASES = ["vdp_gpode_rk4", "vdp_gpode_dopri5", "vdp_shooting_rk4", "vdp_shooting_dopri5", "mocap_gpode_rk4",
         "mocap_shooting_rk4", "d3_shooting_rk4"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_port_reproduces_reference(name):
    g = load_golden(name)
    kind, solver, extra = g['meta'][0], g['meta'][1], ast.literal_eval(g['meta'][3])
    p = {k: v.clone().requires_grad_(True) for k, v in g['p'].items()}
    proj = None
    if g['proj'] is not None:
        comp = g['proj']
        proj = lambda x: torch.einsum('ntl,ld->ntd', x, comp)
    if kind == "gpode":
        r = O.elbo_gpode(p, g['ys'], g['ts'], g['draws'], method=solver, project=proj, **extra)
    else:
        r = O.elbo_shooting(p, g['ys'], g['ts'], g['draws'], method=solver, project=proj)
    r['loss'].backward()
    ref, f64 = g['ref'], g['f64']
    assert_parity(name + " loss", r['loss'], ref['loss'], f64['loss'], TOL_GRAD)
    for k, v in p.items():
        if v.grad is not None and ('grad_' + k) in ref:
            assert_parity(name + " grad " + k, v.grad, ref['grad_' + k], f64['grad_' + k], TOL_GRAD)
    c = r['cache']
    assert relerr(c['rff_omega'], ref['cache_omega']) <= 1e-6
    assert relerr(c['nu'], ref['cache_nu']) <= 1e-4  # same LAPACK path; ill-conditioned but deterministic


@pytest.mark.parametrize("name", CASES)
def test_oracle_flow_reproduces_reference(name):
    g = load_golden(name)
    solver = g['meta'][1]
    gp = O.gp_params(g['p'])
    d = g['draws']
    c = O.build_cache(gp['Z'], gp['Um'], gp['Us_sqrt'], gp['ell'], gp['var'], d['w'], d['eps_omega'], d['phase_u'],
                      d['eps_u'])
    xs = O.flow_forward(g['ref']['traj_in'], g['ref']['traj_grid'], gp, c, method=solver)
    assert_parity(name, xs, g['ref']['traj_out'], g['f64']['traj_out'], TOL_TRAJ)


def test_closed_form_equals_reference_formula():
    g = load_golden("vdp_shooting_rk4")
    gp = O.gp_params(O.cast(g['p'], torch.float64))
    d = O.cast(g['draws'], torch.float64)
    c = O.build_cache(gp['Z'], gp['Um'], gp['Us_sqrt'], gp['ell'], gp['var'], d['w'], d['eps_omega'], d['phase_u'],
                      d['eps_u'])
    x = g['ref']['probe_x'].double()
    a = O.vf_forward(x, gp['Z'], gp['ell'], gp['var'], c)
    b = O.vf_closed_form(x, gp['Z'], gp['ell'], gp['var'], c['rff_omega'], c['rff_phase'], c['rff_weights'], c['nu'])
    assert relerr(a, b) <= 1e-10


def test_rk4_is_three_eighths_rule_and_dopri5_converges():
    """Known-answer checks of the restated torchdiffeq: dy/dt = -y."""
    f = lambda t, y: -y
    y0 = torch.ones(1, 1, dtype=torch.float64)
    t = torch.tensor([0.0, 0.5], dtype=torch.float64)
    h = 0.5
    k1 = -1.0
    k2 = -(1 + h * k1 / 3)
    k3 = -(1 + h * (k2 - k1 / 3))
    k4 = -(1 + h * (k1 - k2 + k3))
    expect = 1 + h * (k1 + 3 * (k2 + k3) + k4) / 8
    got = O.odeint(f, y0, t, method='rk4')[-1].item()
    assert abs(got - expect) < 1e-15
    tt = torch.linspace(0, 2, 5, dtype=torch.float64)
    sol = O.odeint(f, y0, tt, method='dopri5', rtol=1e-8, atol=1e-8)[:, 0, 0]
    assert torch.allclose(sol, torch.exp(-tt), atol=1e-6)
    # decreasing grid: out[0] == y0 and the flow is reversed
    back = O.odeint(f, y0, torch.tensor([0.0, -0.5], dtype=torch.float64), method='rk4')[-1].item()
    assert back > 1.0
