"""TEST INFRASTRUCTURE ONLY -- pins the oracle port against the real reference and writes ``tests/golden/*.npz``.

Run in the build container (needs ``/root/reference``):

    python oracle/pin_against_reference.py            # check + (re)write fixtures
    python oracle/pin_against_reference.py --check    # check only

For every case it runs (1) the UNMODIFIED reference modules (``oracle/reference_harness.py``: reference code +
restated torchdiffeq + injected draws), (2) the oracle port in float32 and (3) the oracle port in float64, and
stores inputs plus the reference's outputs (cache tensors, f(x), trajectories, ELBO terms, every parameter
gradient) and the float64 arbiter values. The GPU box has no ``/root/reference``: the ``-m gpu`` tests compare the
CUDA path with these committed fixtures and with the oracle port.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gpode_oracle as O  # noqa: E402
import reference_harness as H  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (kind, make_problem kwargs, solver, extra)
    "vdp_gpode_rk4": ("gpode", dict(D=2, M=16, S=256, N=1, T=25, seed=121), "rk4", dict(ts_dense_scale=4)),
    "vdp_gpode_dopri5": ("gpode", dict(D=2, M=16, S=256, N=1, T=25, seed=122), "dopri5", dict(ts_dense_scale=4)),
    "vdp_shooting_rk4": ("shooting", dict(D=2, M=16, S=256, N=1, T=25, S_mc=5, seed=123), "rk4", {}),
    "vdp_shooting_dopri5": ("shooting", dict(D=2, M=16, S=256, N=1, T=25, S_mc=5, seed=124), "dopri5", {}),
    "mocap_gpode_rk4": ("gpode", dict(D=5, M=100, S=256, N=2, T=20, D_obs=50, dt=0.01, ell0=1.25, seed=125),
                        "rk4", dict(ts_dense_scale=2)),
    "mocap_shooting_rk4": ("shooting", dict(D=5, M=100, S=256, N=2, T=20, S_mc=3, D_obs=50, dt=0.01, ell0=1.25,
                                            seed=126), "rk4", {}),
    "d3_shooting_rk4": ("shooting", dict(D=3, M=24, S=64, N=2, T=7, S_mc=2, seed=127), "rk4", {}),
}


def _np(tree):
    return {k: v.detach().cpu().numpy() for k, v in tree.items() if torch.is_tensor(v)}


def run_reference(kind, p, ys, ts, draws, proj, S, solver, extra, traj_in, traj_grid):
    mods = H._import_reference()
    if kind == "gpode":
        model = H.build_reference_gpode(mods, p, ys, S, solver=solver, project=proj, **extra)
        with H.injected_draws(mods, draws, n_caches=1, mvn_order=("eps_x0",)):
            loss, terms = H.reference_gpode_loss(model, ys, ts)
    else:
        model = H.build_reference_shooting(mods, p, ys, S, solver=solver, project=proj)
        with H.injected_draws(mods, draws, n_caches=1, mvn_order=("eps_x0", "eps_states")):
            loss, terms = H.reference_shooting_loss(model, ys, ts, num_samples=draws['eps_x0'].shape[0])
    loss.backward()
    gp = model.flow.odefunc.diffeq
    out = dict(loss=loss.detach(), nfe=torch.tensor(model.flow.num_evals()))
    out.update({"term_" + k: v.detach() for k, v in terms.items()})
    out.update({"grad_" + k: v for k, v in H.reference_grads(model, kind).items()})
    # the cache of this ELBO evaluation and f(x) on fixed probe points with it (src/core/dsvgp.py:172-197)
    D = gp.D_in
    xp = torch.tensor(np.random.default_rng(7).normal(size=(64, D)) * 1.5, dtype=ys.dtype)
    with torch.no_grad():
        out.update(cache_omega=gp.rff_omega.detach(), cache_phase=gp.rff_phase.detach(),
                   cache_w=gp.rff_weights.detach(), cache_nu=gp.nu.detach(), probe_x=xp, probe_f=gp(None, xp))
        # Flow.forward (src/core/flow.py:60-90) on fixed initial states and step grid with the same injected cache
        with H.injected_draws(mods, draws, n_caches=1):
            out.update(traj_in=traj_in, traj_grid=traj_grid, traj_out=model.flow(traj_in, traj_grid))
    return out


def run_oracle(kind, p, ys, ts, draws, proj, solver, extra, dtype):
    p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in p.items()}
    ys_, ts_, draws_ = ys.to(dtype), ts.to(dtype), O.cast(draws, dtype)
    stats = {}
    if kind == "gpode":
        r = O.elbo_gpode(p, ys_, ts_, draws_, method=solver, project=proj, stats=stats, **extra)
    else:
        r = O.elbo_shooting(p, ys_, ts_, draws_, method=solver, project=proj, stats=stats)
    r['loss'].backward()
    out = dict(loss=r['loss'].detach(), nfe=torch.tensor(float(stats['nfe'])))
    for k in ("observ_loglik", "init_state_kl", "inducing_kl", "constraint_loglik", "state_entropy"):
        if k in r:
            out["term_" + k] = r[k].detach().reshape(())
    for k, v in p.items():
        if v.grad is not None:
            out["grad_" + k] = v.grad.detach()
    gp = O.gp_params({k: v.detach() for k, v in p.items()})
    c = r['cache']
    xp = torch.tensor(np.random.default_rng(7).normal(size=(64, ys.shape[0] * 0 + gp['Z'].shape[1])) * 1.5,
                      dtype=torch.float32).to(dtype)
    with torch.no_grad():
        cd = {k: v.detach() for k, v in c.items()}
        out.update(cache_omega=cd['rff_omega'], cache_phase=cd['rff_phase'], cache_w=cd['rff_weights'],
                   cache_nu=cd['nu'], probe_x=xp, probe_f=O.vf_forward(xp, gp['Z'], gp['ell'], gp['var'], cd),
                   probe_f_closed=O.vf_closed_form(xp, gp['Z'], gp['ell'], gp['var'], cd['rff_omega'],
                                                   cd['rff_phase'], cd['rff_weights'], cd['nu']))
        if kind == "gpode":
            out['traj_xs'] = r['xs'].detach()
            out['traj_x0'] = r['x0'].detach()
        else:
            out['traj_pred'] = r['pred'].detach()
            out['traj_ss'] = r['ss'].detach()
    return out


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLDEN, exist_ok=True)
    worst = 0.0
    for name, (kind, kw, solver, extra) in CASES.items():
        if args.only and args.only != name:
            continue
        p, ys, ts, draws, proj = O.make_problem(**kw)
        o32 = run_oracle(kind, p, ys, ts, draws, proj, solver, extra, torch.float32)
        if kind == "gpode":
            traj_in = o32.pop('traj_x0')
            traj_grid = O.compute_ts_dense(O.insert_zero_t0(ts), extra['ts_dense_scale'])
        else:
            traj_in, traj_grid = o32['traj_ss'].reshape(-1, kw['D']), ts[:2]
        ref = run_reference(kind, p, ys, ts, draws, proj, kw['S'], solver, extra, traj_in.float(), traj_grid)
        gp32 = O.gp_params(p)
        c32 = O.build_cache(gp32['Z'], gp32['Um'], gp32['Us_sqrt'], gp32['ell'], gp32['var'], draws['w'],
                            draws['eps_omega'], draws['phase_u'], draws['eps_u'])
        o32.update(traj_in=traj_in, traj_grid=traj_grid,
                   traj_out=O.flow_forward(traj_in, traj_grid, gp32, c32, method=solver))
        proj64 = None
        if proj is not None:
            comp64 = proj.components.double()
            proj64 = lambda x: torch.einsum('ntl,ld->ntd', x, comp64)
        o64 = run_oracle(kind, p, ys, ts, draws, proj64, solver, extra, torch.float64)
        o64.pop('traj_x0', None)
        gp64, d64 = O.gp_params(O.cast(p, torch.float64)), O.cast(draws, torch.float64)
        c64 = O.build_cache(gp64['Z'], gp64['Um'], gp64['Us_sqrt'], gp64['ell'], gp64['var'], d64['w'],
                            d64['eps_omega'], d64['phase_u'], d64['eps_u'])
        o64.update(traj_in=traj_in.double(), traj_grid=traj_grid.double(),
                   traj_out=O.flow_forward(traj_in.double(), traj_grid.double(), gp64, c64, method=solver))
        print("== %s  (loss ref %.8f  oracle32 %.8f  oracle64 %.8f, nfe ref %d oracle %d)" % (
            name, ref['loss'], o32['loss'], o64['loss'], ref['nfe'], o32['nfe']))
        # the float64 ARBITER is pinned too: the unmodified reference evaluated in float64 (its dtype singleton switched
        # at run time, reference_harness.reference_in_float64) must agree with the port's float64 run to 1e-6 (measured: 1e-13 shooting, 1e-8 plain GPODE)
        if solver == "rk4":  # dopri5: accept/reject sequences of two float64 runs are identical, but keep this cheap
            with H.reference_in_float64():
                d64r = O.cast(draws, torch.float64)
                p64r = {k: v.double() for k, v in p.items()}
                r64 = run_reference(kind, p64r, ys.double(), ts.double(), d64r, proj, kw['S'], solver, extra,
                                    traj_in.double(), traj_grid.double())
            e_arb = max(relerr(r64[k], o64[k]) for k in r64 if k.startswith("grad_") or k == "loss")
            print("   float64 arbiter: port64-vs-reference64 worst %.2e%s" % (e_arb, "" if e_arb <= 1e-6 else "   <-- MISMATCH"))
            if e_arb > 1e-6:
                worst = max(worst, e_arb)
        for k in sorted(ref):
            if k in ("probe_x", "nfe", "traj_in", "traj_grid"):
                continue
            e32, e64 = relerr(o32[k], ref[k]), relerr(ref[k], o64[k])
            flag = ""
            # the port must reproduce the reference's float32 path to round-off: compare with the same yardstick
            # the reference itself achieves against float64
            tol = max(1e-4 if k.startswith('grad_') else 1e-5, 3 * e64)  # BASELINE.json gates, fp64-arbitrated
            if e32 > tol:
                flag = "   <-- MISMATCH"
                worst = max(worst, e32)
            print("   %-34s port32-vs-ref %.2e   ref-vs-fp64 %.2e%s" % (k, e32, e64, flag))
        if not args.check:
            blob = {}
            blob.update({"in_p_" + k: v for k, v in _np(p).items()})
            blob.update({"in_draw_" + k: v for k, v in _np(draws).items()})
            blob.update(in_ys=ys.numpy(), in_ts=ts.numpy())
            if proj is not None:
                blob['in_proj_components'] = proj.components.numpy()
            blob.update({"ref_" + k: v for k, v in _np(ref).items()})
            blob.update({"f64_" + k: v for k, v in _np(o64).items()})
            blob['meta'] = np.array([kind, solver, repr(kw), repr(extra)])
            np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **blob)
    if worst > 0:
        print("PIN FAILED: worst mismatch %.3e" % worst)
        sys.exit(1)
    print("oracle port pinned against the reference on %d cases" % len(CASES))


if __name__ == "__main__":
    main()
