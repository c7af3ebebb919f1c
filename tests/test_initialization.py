"""Data-driven initialisation against fixtures written by the UNMODIFIED reference
(oracle/make_init_goldens.py): k-means + GP-regression init of the inducing variables (host logic, runs anywhere) and
the backward-in-time initial-state solve through the batched n_sets integrator (-m gpu)."""
import numpy as np
import pytest
import torch

from util import _Queue, _set, load_golden, relerr

CASES = ["init_gpode_dopri5", "init_gpode_rk4", "init_shooting_rk4"]


def _raw(name):
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    return {k: v for k, v in np.load(path, allow_pickle=False).items()}


def _model(g, dev):
    from gaussian_process_odes_b200 import builders
    kind, solver, S = [str(v) for v in g["meta"]]
    p = {k[5:]: torch.tensor(v) for k, v in g.items() if k.startswith("in_p_")}
    N, T, D = g["in_ys"].shape
    M = p["inducing_loc"].shape[0]
    if kind == "gpode":
        model = builders.build_gpode(N, T, D, num_inducing=M, num_features=int(S), solver=solver, ts_dense_scale=4)
    else:
        model = builders.build_gpode_shooting(N, T, D, num_inducing=M, num_features=int(S), solver=solver)
    model = model.to(dev)
    gp = model.flow.odefunc.diffeq
    _set(gp.inducing_loc.optvar, p['inducing_loc'])
    _set(gp.Um.optvar, p['Um'])
    _set(gp.Us_sqrt.optvar, p['Us_sqrt_packed'])
    _set(gp.kern.unconstrained_lengthscales, p['unconstrained_lengthscales'])
    _set(gp.kern.unconstrained_variance, p['unconstrained_variance'])
    return model, kind


@pytest.mark.parametrize("name", CASES[:1] + CASES[2:])
def test_initialize_inducing_matches_reference(name):
    from gaussian_process_odes_b200 import initialization
    g = _raw(name)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    model, _ = _model(g, dev)
    np.random.seed(int(g["in_seed"]))
    initialization.initialize_inducing(model, g["in_ys"], ts_max=float(g["in_ts"][-1]), data_noise=1e-1)
    gp = model.flow.odefunc.diffeq
    # same host generator, same order (subset choice, then scipy's k-means): identical centres
    assert relerr(gp.inducing_loc.optvar.detach().cpu().double(), torch.tensor(g["ref_inducing_loc"])) <= 1e-6
    # whitened inducing means: float64 here vs the reference's float32 factorisations
    assert relerr(gp.Um.optvar.detach().cpu().double(), torch.tensor(g["ref_Um"])) <= 1e-4
    assert gp.Um.optvar.dtype == torch.float32 and gp.inducing_loc.optvar.dtype == torch.float32


def test_noisevar_and_kernel_initialisers():
    from gaussian_process_odes_b200 import builders, initialization
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    model = builders.build_gpode(2, 5, 3, num_inducing=4, num_features=8, D_obs=3).to(dev)
    initialization.initialize_noisevar(model, 0.07)
    assert torch.allclose(model.likelihood.variance, torch.full((3,), 0.07, device=dev), rtol=1e-5)
    initialization.initialize_and_fix_kernel_parameters(model, 1.25, 0.5, fix=True)
    k = model.flow.odefunc.diffeq.kern
    assert torch.allclose(k.lengthscales, torch.full((3, 3), 1.25, device=dev), rtol=1e-5)
    assert torch.allclose(k.variance, torch.full((3,), 0.5, device=dev), rtol=1e-5)
    assert not k.unconstrained_lengthscales.requires_grad and not k.unconstrained_variance.requires_grad


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_initialize_latents_matches_reference(name):
    """n backward-in-time solves, one injected GP draw each, in ONE n_sets launch == the reference's Python loop."""
    from gaussian_process_odes_b200 import initialization
    from gaussian_process_odes_b200.core import dsvgp, kernels
    g = _raw(name)
    model, kind = _model(g, "cuda")
    gp = model.flow.odefunc.diffeq
    _set(gp.inducing_loc.optvar, torch.tensor(g["ref_inducing_loc"]).float())
    _set(gp.Um.optvar, torch.tensor(g["ref_Um"]).float())
    s = {k[7:]: torch.tensor(v) for k, v in g.items() if k.startswith("in_set_")}
    n = int(g["in_n_samples"])
    saved = (dsvgp.sample_normal, dsvgp.sample_uniform, kernels.sample_normal)
    dsvgp.sample_normal = _Queue([x for q in range(n) for x in (s['w'][q], s['eps_u'][q])], "dsvgp.normal")
    dsvgp.sample_uniform = _Queue([s['phase_u'][q] for q in range(n)], "dsvgp.uniform")
    kernels.sample_normal = _Queue([s['eps_omega'][q] for q in range(n)], "kernels.normal")
    try:
        initialization.initialize_latents_with_data(model, g["in_ys"], g["in_ts"], num_samples=n)
    finally:
        dsvgp.sample_normal, dsvgp.sample_uniform, kernels.sample_normal = saved
    x0d = model.x0_distribution if kind == "gpode" else model.state_distribution.x0
    assert relerr(x0d.param_mean.optvar.detach().cpu(), torch.tensor(g["ref_x0_mean"])) <= 1e-4
    if kind != "gpode":
        assert torch.equal(model.state_distribution.param_mean.optvar.detach().cpu(), torch.tensor(g["ref_state_mean"]))
