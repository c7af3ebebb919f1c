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


def test_staged_reference_archive_reproduces_the_golden_loss(tmp_path):
    """The CPU arm of bench.py (`--impl reference`, `cpu_baseline`) imports the UNMODIFIED reference from the archive
    `oracle/_ref/gpode_reference_src.zip` that `__graft_entry__.build()` stages (oracle/stage_reference.py). Run in a
    fresh interpreter with the archive as the only source of the reference: it must import (zipimport, implicit namespace
    packages), stay on the CPU whatever torch says about CUDA, and reproduce the committed golden loss bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    archive = os.path.join(root, "oracle", "_ref", "gpode_reference_src.zip")
    if not os.path.isfile(archive):
        if not os.path.isdir("/root/reference/src"):
            pytest.skip("no staged reference archive and no reference tree on this machine")
        sys.path.insert(0, os.path.join(root, "oracle"))
        import stage_reference
        assert stage_reference.stage() == archive
    script = tmp_path / "run_ref.py"
    script.write_text('''
import os, sys
root = %r
sys.path[:0] = [os.path.join(root, "oracle"), os.path.join(root, "tests")]
import torch
torch.cuda.is_available = lambda: True   # what the reference's device singleton sees on the GPU box
import reference_harness as H
from util import load_golden
assert H.source() == "archive", H.source()
g = load_golden("vdp_shooting_rk4")
mods = H._import_reference()
assert ".zip" in mods["dsvgp"].__file__, mods["dsvgp"].__file__
model = H.build_reference_shooting(mods, g["p"], g["ys"], 256, solver="rk4")
with H.injected_draws(mods, g["draws"], n_caches=1, mvn_order=("eps_x0", "eps_states")):
    loss, _ = H.reference_shooting_loss(model, g["ys"], g["ts"], num_samples=g["draws"]["eps_x0"].shape[0])
loss.backward()
print("LOSS %%.9e REF %%.9e" %% (float(loss), float(g["ref"]["loss"])))
assert abs(float(loss) - float(g["ref"]["loss"])) <= 1e-6 * abs(float(g["ref"]["loss"]))
''' % root)
    env = dict(os.environ, GPODE_REFERENCE_ROOT=archive, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
