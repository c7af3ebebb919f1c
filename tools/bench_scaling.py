#!/usr/bin/env python
"""BASELINE.json configs 2 and 5 at 1 / 2 / 4 / 8 GPUs (config 4 is bench.py itself). One process per GPU:

    python tools/bench_scaling.py                                              # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_scaling.py

* config 2, VDP shooting (train_vdp_gpode_shooting.py defaults: D=2, M=16, S=256, N=1, T=25, S_mc=5 -> 125 segments):
  ONE sequence, so the TIME axis is sharded (distributed.enable_time_sharding: each rank samples and integrates its
  slice plus one halo state) -- or, for comparison, the segment ROWS (distributed.enable_row_sharding; the constraint's neighbour state is
  the halo row of the replicated sample tensor) and the flattened gradient of every parameter is all-reduced once.
  ELBO fwd+bwd step ms; also a long variant (T = 200 000 -> 10^6 segments) of the same model.
* config 5, scaling sweep: D in {2, 5, 16, 64}, 10^6 rows in total (strong scaling: 10^6 / N per GPU), S=256, M=16 (D=2)
  else 100, one RK4 step of h = 0.01 forward + discrete adjoint (gradients w.r.t. x0, Z, lengthscales, variances, nu)
  with the shared-parameter gradient all-reduced; vector-field evaluations per second = 4 x rows / step time.
Timing: CUDA events on each rank, barrier on both sides, max over ranks, median of the repetitions."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gpode_oracle as O  # noqa: E402
from util import build_product_model  # noqa: E402


def timed(fn, world, dev, warm=3, reps=10):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out.append(float(ms))
    return float(np.median(out))


def vdp_shooting(rank, world, dev, T, mode="time"):
    from gaussian_process_odes_b200 import distributed
    kw = dict(D=2, M=16, S=256, N=1, T=T, S_mc=5)
    p, ys, ts, draws, _ = O.make_problem(seed=121, **kw)
    model = build_product_model("shooting", p, ys, kw['S'], "rk4")
    # mode "time": every rank owns a slice of the time axis (state-distribution work shards too); "rows": contiguous row
    # blocks with the whole state distribution replicated (round 2, first version: flat from 1 to 8 GPUs)
    if mode == "time":
        distributed.enable_time_sharding(model, rank, world)
    else:
        distributed.enable_row_sharding(model, rank, world)
    ys, ts = ys.to(dev), ts.to(dev)

    def loss_fn():
        if mode == "time":
            return distributed.time_sharded_shooting_loss(model, ys, ts, kw['S_mc'], world)
        return distributed.row_sharded_shooting_loss(model, ys, ts, kw['S_mc'], world)

    def eager_step():
        distributed.seed_ranks(7, rank, same_states=True)   # every rank draws the same states and the same GP function
        model.zero_grad(set_to_none=True)
        loss_fn().backward()
        distributed.allreduce_all_grads(model)
    ms_eager = timed(eager_step, world, dev)
    # An eager step is ~150 host-issued launches: 3-4 ms of HOST time whatever the segment count, so it cannot scale.
    # The library's form for such steps is the CUDA-graph replay (graphs.GraphedStep, as bench.py uses): ELBO forward +
    # backward captured once, the gradient all-reduce outside the graph. Every rank's generators are seeded identically
    # before the capture and consume the same number of draws per step, so the ranks keep drawing the same noise.
    ms = ms_eager
    graphed = False
    try:
        from gaussian_process_odes_b200 import graphs
        distributed.seed_ranks(7, rank, same_states=True)
        model.zero_grad(set_to_none=True)
        gstep = graphs.GraphedStep(model, loss_fn)

        def graph_step():
            gstep()
            distributed.allreduce_all_grads(model)
        ms = timed(graph_step, world, dev)
        graphed = True
    except Exception as exc:   # capture is an optimisation: report the eager number
        if rank == 0:
            print("CUDA-graph capture failed (%r): eager timing reported" % (exc,), file=sys.stderr)
    rows = kw['S_mc'] * kw['N'] * T
    name = ("vdp_shooting" if T == 25 else "vdp_shooting_long") + ("" if mode == "time" else "_row_sharded")
    return dict(config=name, n_gpus=world, segments_total=rows,
                parallelism=("time axis sharded x%d (each rank samples and integrates its slice + one halo state; one "
                             "all-reduce of all gradients)" if mode == "time" else
                             "segment rows sharded x%d (replicated parameters, one all-reduce of all gradients)") % world,
                cuda_graph=graphed, elbo_fwd_bwd_ms=ms, elbo_fwd_bwd_ms_eager=ms_eager,
                evals_per_s=4 * rows / (ms * 1e-3))


def sweep(rank, world, dev, D, M, B_total):
    from gaussian_process_odes_b200 import ops
    from gaussian_process_odes_b200.distributed import shard_range
    S = 256
    p, ys, ts, draws, _ = O.make_problem(D=D, M=M, S=S, N=1, T=4, seed=5)
    gp = O.gp_params(p)
    omega = draws['eps_omega'] / gp['ell'].T.unsqueeze(1)
    nu = torch.tensor(np.random.default_rng(1).normal(size=(D, M)) * 0.1, dtype=torch.float32)
    args = [a.to(dev).contiguous() for a in (gp['Z'], gp['ell'], gp['var'], nu, omega, draws['phase_u'] * 2 * np.pi,
                                             draws['w'])]
    for a in args[:4]:
        a.requires_grad_(True)
    lo, hi = shard_range(B_total, rank, world)
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(B_total, D, generator=g)[lo:hi].to(dev).requires_grad_(True)
    cot = torch.randn(2, B_total, D, generator=g)[:, lo:hi].to(dev).contiguous()
    tg = (torch.arange(2, dtype=torch.float32) * 0.01).to(dev)

    def step():
        for a in args[:4]:
            a.grad = None
        x.grad = None
        xs = ops.rk4_integrate(x, tg, *args)
        xs.backward(cot)
        if world > 1:
            flat = torch.cat([a.grad.reshape(-1) for a in args[:4]])
            dist.all_reduce(flat)
    ms = timed(step, world, dev, warm=2, reps=5 if D > 8 else 10)
    fv = D * (S * (2 * D + 4) + M * (3 * D + 4))
    return dict(config="sweep", D=D, M=M, S=S, n_gpus=world, rows_total=B_total, rows_per_gpu=hi - lo,
                parallelism="rows sharded x%d, one all-reduce of the shared-parameter gradient" % world,
                rk4_step_fwd_bwd_ms=ms, evals_per_s=4 * B_total / (ms * 1e-3),
                algorithmic_tflops_per_gpu=(hi - lo) * 12 * fv / (ms * 1e-3) / 1e12,
                flops_note="4 F_vf forward + 8 F_vf adjoint per row-step (SURVEY 8d)")


def main():
    from gaussian_process_odes_b200 import distributed
    rank, world, local = distributed.init_from_env()
    dev = torch.device("cuda", local)
    only = os.environ.get("SCALING_ONLY", "")
    res = []

    def emit(r):
        res.append(r)
        if rank == 0:
            print(json.dumps(r), flush=True)
    if not only or "vdp" in only:
        emit(vdp_shooting(rank, world, dev, 25))
        emit(vdp_shooting(rank, world, dev, 200000))
        emit(vdp_shooting(rank, world, dev, 200000, mode="rows"))
    if not only or "sweep" in only:
        for D, M in ((2, 16), (5, 100), (16, 100), (64, 100)):
            emit(sweep(rank, world, dev, D, M, 1000000 if D < 64 else 200000))
    if rank == 0 and len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
