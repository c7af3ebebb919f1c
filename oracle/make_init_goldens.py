"""TEST INFRASTRUCTURE ONLY -- golden fixtures for the reference's data-driven initialisation (SURVEY.md section 8f
item 4): ``initialize_inducing`` and ``initialize_latents_with_data`` of ``src/gpode/model_initialization.py`` and
``src/gpode_shooting/model_initialization.py``.

Runs only the UNMODIFIED reference (``oracle/reference_harness.py``; needs ``/root/reference``) on a small synthetic
Van-der-Pol-like data set. ``initialize_inducing`` consumes numpy's global generator (observation subset, scipy
k-means) -- seeded; the backward-in-time integrations draw one GP function per sample -- injected, a different draw
per sample. Stores inputs + the reference's outputs in ``tests/golden/init_*.npz``.

    python oracle/make_init_goldens.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gpode_oracle as O  # noqa: E402
import reference_harness as H  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


class _PerSampleDraws:
    """build_cache call q of the reference consumes draw q (w, omega, phase, epsilon -- dsvgp.py:100-103,83)."""

    def __init__(self, mods, sets):
        self.m, self.s = mods, sets

    def __enter__(self):
        m, s = self.m, self.s
        n = s['w'].shape[0]
        self.saved = (m['dsvgp'].sample_normal, m['dsvgp'].sample_uniform, m['kernels'].sample_normal)
        m['dsvgp'].sample_normal = H._Queue([x for q in range(n) for x in (s['w'][q], s['eps_u'][q])], "dsvgp.normal")
        m['dsvgp'].sample_uniform = H._Queue([s['phase_u'][q] for q in range(n)], "dsvgp.uniform")
        m['kernels'].sample_normal = H._Queue([s['eps_omega'][q] for q in range(n)], "kernels.normal")
        return self

    def __exit__(self, *exc):
        m = self.m
        m['dsvgp'].sample_normal, m['dsvgp'].sample_uniform, m['kernels'].sample_normal = self.saved
        return False


def run(name, kind, solver, D=2, M=8, S=32, N=3, T=14, n_samples=4, seed=17):
    mods = H._import_reference()
    init = importlib.import_module("src.gpode.model_initialization" if kind == "gpode"
                                   else "src.gpode_shooting.model_initialization")
    p, _, _, _, _ = O.make_problem(D=D, M=M, S=S, N=N, T=T, seed=seed, S_mc=1)
    rng = np.random.default_rng(seed)
    # smooth 2-D orbits + noise, float32 like every tensor on the reference's path
    ts = np.linspace(0.0, 3.0, T).astype(np.float32)
    ph = rng.uniform(0, 2 * np.pi, size=(N, 1))
    amp = rng.uniform(1.0, 2.0, size=(N, 1))
    ys = np.stack([amp * np.cos(ts[None] + ph), -amp * np.sin(ts[None] + ph)], -1)
    if D > 2:
        ys = np.concatenate([ys, rng.normal(size=(N, T, D - 2)) * 0.3], -1)
    ys = (ys + rng.normal(size=ys.shape) * 0.05).astype(np.float32)
    yt = torch.tensor(ys)
    build = H.build_reference_gpode if kind == "gpode" else H.build_reference_shooting
    model = build(mods, p, yt, S, solver=solver)
    np.random.seed(seed)
    init.initialize_inducing(model, ys, ts_max=float(ts[-1]), data_noise=1e-1)
    gp = model.flow.odefunc.diffeq
    Z, Um = gp.inducing_loc.optvar.detach().double().clone(), gp.Um.optvar.detach().double().clone()
    # the reference assigns whatever dtype scipy returned; keep the layer float32 for the integrations below
    gp.inducing_loc.optvar.data = Z.float()
    gp.Um.optvar.data = Um.float()
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    sets = dict(w=t(rng.normal(size=(n_samples, S, D))), eps_omega=t(rng.normal(size=(n_samples, D, S, D))),
                phase_u=t(rng.uniform(size=(n_samples, 1, S, D))), eps_u=t(rng.normal(size=(n_samples, M, D))))
    with _PerSampleDraws(mods, sets):
        init.initialize_latents_with_data(model, ys, ts, num_samples=n_samples)
    x0d = model.x0_distribution if kind == "gpode" else model.state_distribution.x0
    blob = dict(in_ys=ys, in_ts=ts, in_seed=np.array(seed), in_n_samples=np.array(n_samples),
                meta=np.array([kind, solver, str(S)]),
                ref_inducing_loc=Z.numpy(), ref_Um=Um.numpy(), ref_x0_mean=x0d.param_mean.optvar.detach().numpy())
    if kind != "gpode":
        blob["ref_state_mean"] = model.state_distribution.param_mean.optvar.detach().numpy()
    blob.update({"in_p_" + k: v.numpy() for k, v in p.items()})
    blob.update({"in_set_" + k: v.numpy() for k, v in sets.items()})
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **blob)
    print(name, "Z", tuple(Z.shape), "|Um|max %.3f" % Um.abs().max().item(),
          "x0", blob["ref_x0_mean"].round(3).tolist(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    run("init_gpode_dopri5", "gpode", "dopri5")
    run("init_gpode_rk4", "gpode", "rk4")
    run("init_shooting_rk4", "shooting", "rk4")
