"""TEST INFRASTRUCTURE ONLY -- golden fixtures for the reference's NON-DEFAULT layer options (SURVEY.md section 8f
item 4): ``dimwise=False`` (one kernel shared by all output dimensions) and ``q_diag=True`` (diagonal q(u)).

Runs only the UNMODIFIED reference (``oracle/reference_harness.py``; needs ``/root/reference``) with injected draws
and stores inputs + the reference's ELBO, gradients and a probe evaluation of f(x) in ``tests/golden/variant_*.npz``.
There is no oracle-port restatement of these two branches, so the tests compare against the reference's float32
numbers directly.

    python oracle/make_variant_goldens.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gpode_oracle as O  # noqa: E402
import reference_harness as H  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _reference_f64(mods, dimwise, q_diag, D, M, S, N, T, S_mc, state_dict, ys, ts, draws, xp):
    with H.reference_in_float64():
        gp = mods['dsvgp'].DSVGP_Layer(D_in=D, D_out=D, M=M, S=S, dimwise=dimwise, q_diag=q_diag)
        flow = mods['flow'].Flow(diffeq=gp, solver='rk4', use_adjoint=False)
        lik = mods['likelihoods'].Gaussian(ndim=D)
        cons = mods['constraints'].Gaussian(d=1, scale=1e-2, requires_grad=False)
        sd = mods['states'].StateSequenceVariationalFactorizedGaussian(dim_n=N, dim_t=T - 1, dim_d=D)
        model = mods['shooting_models'].UniformSequenceModel(flow=flow, num_observations=N * T * D,
                                                             state_distribution=sd, likelihood=lik, constraint=cons)
        model = model.double()
        model.load_state_dict({k: v.detach().double() for k, v in state_dict.items()})
        d64 = {k: v.double() for k, v in draws.items()}
        with H.injected_draws(mods, d64, mvn_order=("eps_x0", "eps_states")):
            loss, _ = H.reference_shooting_loss(model, ys.double(), ts.double(), num_samples=S_mc)
        assert loss.dtype == torch.float64
        loss.backward()
        with torch.no_grad():
            probe = gp(None, xp.double())
        assert probe.dtype == torch.float64
        grads = {n: p.grad.detach().numpy() for n, p in model.named_parameters() if p.grad is not None}
        assert all(g.dtype == np.float64 for g in grads.values())
        return loss.detach().numpy(), grads, probe.numpy()


def run(name, dimwise, q_diag, D=2, M=16, S=64, N=2, T=9, S_mc=3, seed=31):
    mods = H._import_reference()
    rng = np.random.default_rng(seed)
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    ys = t(rng.normal(size=(N, T, D)) * 1.5)
    ts = t(np.linspace(0.0, 2.0, T))
    gp = mods['dsvgp'].DSVGP_Layer(D_in=D, D_out=D, M=M, S=S, dimwise=dimwise, q_diag=q_diag)
    flow = mods['flow'].Flow(diffeq=gp, solver='rk4', use_adjoint=False)
    lik = mods['likelihoods'].Gaussian(ndim=D)
    cons = mods['constraints'].Gaussian(d=1, scale=1e-2, requires_grad=False)
    sd = mods['states'].StateSequenceVariationalFactorizedGaussian(dim_n=N, dim_t=T - 1, dim_d=D)
    model = mods['shooting_models'].UniformSequenceModel(flow=flow, num_observations=N * T * D, state_distribution=sd,
                                                         likelihood=lik, constraint=cons)
    with torch.no_grad():  # perturb the default init so that every gradient is exercised
        gp.inducing_loc.optvar.copy_(t(rng.normal(size=(M, D)) * 1.5))
        gp.Um.optvar.copy_(t(rng.normal(size=(M, D)) * 0.1))
        gp.Us_sqrt.optvar.add_(t(rng.normal(size=tuple(gp.Us_sqrt.optvar.shape)) * 1e-2))
        gp.kern.unconstrained_lengthscales.add_(t(rng.normal(size=tuple(gp.kern.unconstrained_lengthscales.shape)) * 0.1))
        gp.kern.unconstrained_variance.add_(t(rng.normal(size=tuple(gp.kern.unconstrained_variance.shape)) * 0.1))
        sd.param_mean.optvar.copy_(t(rng.normal(size=(N, T - 1, D))))
        sd.x0.param_mean.optvar.copy_(t(rng.normal(size=(N, D))))
    phase_shape = (1, S, D) if dimwise else (1, S)
    omega_shape = (D, S, D) if dimwise else (D, S)
    draws = dict(w=t(rng.normal(size=(S, D))), eps_omega=t(rng.normal(size=omega_shape)),
                 phase_u=t(rng.uniform(size=phase_shape)), eps_u=t(rng.normal(size=(M, D))),
                 eps_x0=t(rng.normal(size=(S_mc, N, D))), eps_states=t(rng.normal(size=(S_mc, N, T - 1, D))))
    with H.injected_draws(mods, draws, mvn_order=("eps_x0", "eps_states")):
        loss, terms = H.reference_shooting_loss(model, ys, ts, num_samples=S_mc)
    loss.backward()
    xp = t(np.random.default_rng(7).normal(size=(40, D)) * 1.5)
    with torch.no_grad():
        probe = gp(None, xp)
    blob = dict(in_ys=ys.numpy(), in_ts=ts.numpy(), probe_x=xp.numpy(), ref_probe_f=probe.numpy(),
                ref_loss=loss.detach().numpy(), meta=np.array([str(dimwise), str(q_diag), str(S), str(S_mc)]))
    blob.update({"in_draw_" + k: v.numpy() for k, v in draws.items()})
    blob.update({"in_sd_" + k: v.detach().numpy() for k, v in model.state_dict().items()})
    blob.update({"ref_grad_" + n: p.grad.detach().numpy() for n, p in model.named_parameters() if p.grad is not None})
    blob.update({"ref_term_" + k: v.detach().numpy().reshape(()) for k, v in terms.items()})
    if dimwise:
        # float64 arbiter from the oracle port (which restates the dimwise branch; q_diag via a diagonal embedding)
        sdm = {k: v.detach().double() for k, v in model.state_dict().items()}
        pre = "flow.odefunc.diffeq."
        pp = dict(inducing_loc=sdm[pre + "inducing_loc.optvar"], Um=sdm[pre + "Um.optvar"],
                  unconstrained_lengthscales=sdm[pre + "kern.unconstrained_lengthscales"],
                  unconstrained_variance=sdm[pre + "kern.unconstrained_variance"],
                  x0_mean=sdm["state_distribution.x0.param_mean.optvar"],
                  x0_lchol_packed=sdm["state_distribution.x0.param_lchol.optvar"],
                  state_mean=sdm["state_distribution.param_mean.optvar"],
                  state_lchol_packed=sdm["state_distribution.param_lchol.optvar"],
                  lik_unconstrained_variance=sdm["likelihood.unconstrained_variance"],
                  constraint_unconstrained_scale=sdm["constraint.unconstrained_scale"])
        pp["Us_sqrt_diag_unconstrained" if q_diag else "Us_sqrt_packed"] = sdm[pre + "Us_sqrt.optvar"]
        pp = {k: v.clone().requires_grad_(True) for k, v in pp.items()}
        r = O.elbo_shooting(pp, ys.double(), ts.double(), O.cast(draws, torch.float64), method="rk4")
        r["loss"].backward()
        names = {"inducing_loc": pre + "inducing_loc.optvar", "Um": pre + "Um.optvar",
                 "unconstrained_lengthscales": pre + "kern.unconstrained_lengthscales",
                 "unconstrained_variance": pre + "kern.unconstrained_variance",
                 "state_mean": "state_distribution.param_mean.optvar",
                 "x0_mean": "state_distribution.x0.param_mean.optvar",
                 "lik_unconstrained_variance": "likelihood.unconstrained_variance",
                 ("Us_sqrt_diag_unconstrained" if q_diag else "Us_sqrt_packed"): pre + "Us_sqrt.optvar"}
        blob["f64_loss"] = r["loss"].detach().numpy()
        for k, n in names.items():
            blob["f64_grad_" + n] = pp[k].grad.numpy()
        print("   fp64 arbiter: loss", float(r["loss"]), " ref32-vs-fp64 lengthscale grad",
              float(np.abs(blob["ref_grad_" + pre + "kern.unconstrained_lengthscales"]
                           - blob["f64_grad_" + pre + "kern.unconstrained_lengthscales"]).max()
                    / np.abs(blob["f64_grad_" + pre + "kern.unconstrained_lengthscales"]).max()))
    # float64 arbiter from the reference itself (every branch, including dimwise=False)
    l64, g64, p64 = _reference_f64(mods, dimwise, q_diag, D, M, S, N, T, S_mc, model.state_dict(), ys, ts, draws, xp)
    blob["r64_loss"], blob["r64_probe_f"] = l64, p64
    blob.update({"r64_grad_" + n: g for n, g in g64.items()})
    worst = max(float(np.abs(blob["ref_grad_" + n] - g).max() / (np.abs(g).max() + 1e-300)) for n, g in g64.items())
    print("   reference float64 pass: loss %.10f (float32 %.10f), worst ref32-vs-ref64 gradient error %.2e, probe f %.2e"
          % (float(l64), float(loss), worst, float(np.abs(probe.numpy() - p64).max() / np.abs(p64).max())))
    if "f64_loss" in blob:
        print("   oracle-port float64 vs reference float64: loss %.2e" % abs(float(blob["f64_loss"]) - float(l64)))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **blob)
    print(name, "loss", float(loss), "grads", sorted(k for k in blob if k.startswith("ref_grad_")))


if __name__ == "__main__":
    torch.manual_seed(0)
    np.random.seed(0)
    run("variant_nodimwise", dimwise=False, q_diag=False)
    run("variant_qdiag", dimwise=True, q_diag=True)
    run("variant_nodimwise_qdiag", dimwise=False, q_diag=True, D=3, M=12, S=33)
