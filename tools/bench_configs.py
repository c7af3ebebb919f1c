#!/usr/bin/env python
"""ELBO fwd+bwd step time of the four BASELINE.json model configs and vector-field evals/s of the scaling sweep
(SURVEY.md section 8d), CUDA path next to the oracle port on the host cores.

    python tools/bench_configs.py [--no-cpu] [--out profiles/xyz.json]

Timing: CUDA events, 5 warm-ups, median of 20 (GPU); perf_counter, 2 warm-ups, median of 5 (CPU oracle port)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpode_oracle as O  # noqa: E402
from util import build_product_model  # noqa: E402

CONFIGS = {
    "vdp_gpode":       dict(kind="gpode", D=2, M=16, S=256, N=1, T=25, S_mc=1, ts_dense_scale=4),
    "vdp_shooting":    dict(kind="shooting", D=2, M=16, S=256, N=1, T=25, S_mc=5),
    "mocap_gpode":     dict(kind="gpode", D=5, M=100, S=256, N=6, T=100, S_mc=1, D_obs=50, dt=0.01, ell0=1.25,
                            ts_dense_scale=2),
    "mocap_shooting":  dict(kind="shooting", D=5, M=100, S=256, N=6, T=100, S_mc=5, D_obs=50, dt=0.01, ell0=1.25),
}


def gpu_time(fn, warm=5, reps=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def wall_time(fn, warm=5, reps=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def cpu_time(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def run_config(name, c, solver, do_cpu):
    from gaussian_process_odes_b200 import builders
    kw = {k: v for k, v in c.items() if k not in ("kind", "ts_dense_scale")}
    p, ys, ts, draws, proj = O.make_problem(seed=121, **kw)
    comp = proj.components if proj is not None else None
    model = build_product_model(c["kind"], p, ys, c["S"], solver, ts_dense_scale=c.get("ts_dense_scale", 4), proj=comp)
    ysd, tsd = ys.cuda(), ts.cuda()

    def step():
        model.zero_grad(set_to_none=True)
        if c["kind"] == "gpode":
            loss = builders.compute_loss_gpode(model, ysd, tsd)[0]
        else:
            loss = builders.compute_loss_shooting(model, ysd, tsd, num_samples=c["S_mc"])[0]
        loss.backward()

    def fwd():
        with torch.no_grad():
            if c["kind"] == "gpode":
                builders.compute_loss_gpode(model, ysd, tsd)
            else:
                builders.compute_loss_shooting(model, ysd, tsd, num_samples=c["S_mc"])

    out = dict(config=name, solver=solver, rows=(c["S_mc"] * c["N"] * c["T"] if c["kind"] == "shooting" else c["N"]))
    if True:  # both solvers: dopri5 training is captured with the accepted-step count left on the device
        out["gpu_fwd_bwd_ms"] = gpu_time(step)
        out["gpu_fwd_bwd_wall_ms"] = wall_time(step)
        from gaussian_process_odes_b200 import graphs
        if c["kind"] == "gpode":
            gstep = graphs.GraphedStep(model, lambda: builders.compute_loss_gpode(model, ysd, tsd)[0])
        else:
            gstep = graphs.GraphedStep(model, lambda: builders.compute_loss_shooting(model, ysd, tsd,
                                                                                    num_samples=c["S_mc"])[0])
        out["gpu_fwd_bwd_graphed_ms"] = gpu_time(gstep)
        out["gpu_fwd_bwd_graphed_wall_ms"] = wall_time(gstep)
        if solver == "dopri5":
            out["graphed_dopri5_stats"] = gstep.check()
    out["gpu_fwd_ms"] = gpu_time(fwd)
    out["nfe"] = model.flow.num_evals()
    if do_cpu:
        def cpu_step():
            pp = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
            if c["kind"] == "gpode":
                r = O.elbo_gpode(pp, ys, ts, draws, method=solver, project=proj, ts_dense_scale=c["ts_dense_scale"])
            else:
                r = O.elbo_shooting(pp, ys, ts, draws, method=solver, project=proj)
            r["loss"].backward()
        out["cpu_port_fwd_bwd_ms"] = cpu_time(cpu_step)
        out["cpu_threads"] = torch.get_num_threads()
    return out


def run_prediction(n_samples=64):
    """the notebook's prediction rate (VDP GPODE, dopri5, ts_dense_scale=2, 51-point grid,
    one new GP function draw per sample; reference: 3.54 it/s on an unknown CPU)."""
    from gaussian_process_odes_b200 import builders
    c = dict(D=2, M=16, S=256, N=1, T=51)
    p, ys, ts, draws, _ = O.make_problem(seed=121, **c)
    model = build_product_model("gpode", p, ys, c["S"], "dopri5", ts_dense_scale=2)
    tsd = ts.cuda()
    builders.compute_predictions(model, tsd, eval_sample_size=4, batched=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = builders.compute_predictions(model, tsd, eval_sample_size=n_samples, batched=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = dict(prediction=True, config="vdp_gpode_notebook", solver="dopri5", samples=n_samples,
               it_per_s=n_samples / dt, ms_per_sample=dt / n_samples * 1e3, out_shape=list(out.shape),
               reference_it_per_s=3.54)
    # all draws in one whitening + pack + integrator launch (n_sets path)
    for rng in ("numpy", "device"):
        for ns in (128, 1024):
            builders.compute_predictions(model, tsd, eval_sample_size=ns, rng=rng)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                outb = builders.compute_predictions(model, tsd, eval_sample_size=ns, rng=rng)
            torch.cuda.synchronize()
            dtb = (time.perf_counter() - t0) / 5
            res["batched_%s_%d_it_per_s" % (rng, ns)] = ns / dtb
            res["batched_%s_%d_ms_total" % (rng, ns)] = dtb * 1e3
    res["batched_out_shape"] = list(outb.shape)
    from gaussian_process_odes_b200 import _lib
    _lib.profile_start()
    builders.compute_predictions(model, tsd, eval_sample_size=128, rng="device")
    res["batched_128_kernels_ms"] = {k: round(v[1], 4) for k, v in _lib.profile_stop().items()}
    try:
        from gaussian_process_odes_b200 import graphs
        gp = graphs.GraphedPrediction(model, tsd)
        gp.sample_many(4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out2 = gp.sample_many(n_samples)
        torch.cuda.synchronize()
        dt2 = time.perf_counter() - t0
        res.update(graphed_it_per_s=n_samples / dt2, graphed_ms_per_sample=dt2 / n_samples * 1e3,
                   graphed_out_shape=list(out2.shape))
    except Exception as e:  # cooperative launches may not be capturable on every driver
        res["graphed_error"] = repr(e)[:200]
    return res


def run_sweep(D, M, S, B, K, do_cpu):
    from gaussian_process_odes_b200 import ops
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=5)
    gp = O.gp_params(p)
    d = draws
    omega = d['eps_omega'] / gp['ell'].T.unsqueeze(1)
    nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.1, dtype=torch.float32)
    args = [a.cuda().contiguous() for a in (gp['Z'], gp['ell'], gp['var'], nu, omega, d['phase_u'] * 2 * np.pi, d['w'])]
    x = torch.randn(B, D, device="cuda")
    tg = (torch.arange(K + 1, dtype=torch.float32) * 0.01).cuda()
    reps = 10 if D <= 8 else 3
    with torch.no_grad():
        ms = gpu_time(lambda: ops.rk4_integrate(x, tg, *args), warm=2, reps=reps)
    fv = D * (S * (2 * D + 4) + M * (3 * D + 4))
    evals = B * 4 * K
    out = dict(sweep=True, D=D, M=M, S=S, B=B, rk4_steps=K, ms=ms, evals_per_s=evals / (ms * 1e-3),
               algorithmic_tflops=evals * fv / (ms * 1e-3) / 1e12)
    if K == 32:  # batch-norm dopri5 over [0, 0.32], rtol = atol = 1e-6
        td = torch.tensor([0.0, 0.32], device="cuda")
        with torch.no_grad():
            _, st = ops.dopri5_integrate(x, td, *args)
            ms5 = gpu_time(lambda: ops.dopri5_integrate(x, td, *args), warm=1, reps=max(reps // 2, 2))
        nfe = int(st[0])
        out.update(dopri5_ms=ms5, dopri5_nfe=nfe, dopri5_evals_per_s=B * nfe / (ms5 * 1e-3),
                   dopri5_algorithmic_tflops=B * nfe * fv / (ms5 * 1e-3) / 1e12)
    if do_cpu and B <= 100000:
        c = dict(rff_omega=omega, rff_phase=d['phase_u'] * 2 * np.pi, rff_weights=d['w'], nu=nu.unsqueeze(2))
        xc = x.cpu()
        with torch.no_grad():
            t = cpu_time(lambda: O.vf_forward(xc, gp['Z'], gp['ell'], gp['var'], c), warm=1, reps=3)
        out["cpu_port_vf_evals_per_s"] = B / (t * 1e-3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    res = []
    for name, c in CONFIGS.items():
        if args.only and args.only not in name:
            continue
        for solver in ("rk4", "dopri5"):
            r = run_config(name, c, solver, not args.no_cpu and solver == "rk4")
            print(json.dumps(r), flush=True)
            res.append(r)
    if not args.only or args.only == "prediction":
        r = run_prediction()
        print(json.dumps(r), flush=True)
        res.append(r)
    if not args.only or args.only == "sweep":
        only_d = [int(v) for v in os.environ.get("SWEEP_D", "").split(",") if v]
        for D, M in ((2, 16), (4, 100), (5, 100), (8, 100), (16, 100), (32, 100), (64, 100)):
            if only_d and D not in only_d:
                continue
            shapes = ((10000, 32), (100000, 32), (1000000, 1), (1000000, 32)) if D <= 8 else \
                     ((10000, 32), (100000, 1), (100000, 32))
            for B, K in shapes:
                r = run_sweep(D, M, 256, B, K, not args.no_cpu)
                print(json.dumps(r), flush=True)
                res.append(r)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
