"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun, NCCL). Every rank runs the UNSHARDED problem on
its own GPU through the CUDA path, then the sharded one with the other ranks, and compares loss and gradients.

  A. segment-ROW sharding (distributed.enable_row_sharding): VDP shooting, N = 1 sequence -- the (S_mc, N, T) batch is
     cut into contiguous row blocks across samples and time; all parameters replicated; one all-reduce of every gradient.
  A2. TIME sharding (distributed.enable_time_sharding): the same problem, each rank owns a slice of the time axis -- the
     state-distribution work shards too; the halo state's gradient crosses ranks through the all-reduce.
  B. SEQUENCE sharding (bench.py's mode): MoCap-shaped problem, each rank owns N / world sequences and their variational
     states; one all-reduce of the shared-parameter gradient.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import gpode_oracle as O  # noqa: E402
from util import build_product_model, injected_draws, product_grads, relerr  # noqa: E402


def unsharded(p, ys, ts, draws, proj, kw):
    model = build_product_model("shooting", p, ys, kw['S'], "rk4", proj=None if proj is None else proj.components)
    with injected_draws(draws, mvn_order=("eps_x0", "eps_states")):
        ll, c, e, k0 = model.build_lowerbound_terms(ys.cuda(), ts.cuda(), num_samples=kw['S_mc'])
        loss = -(ll + c + e - k0 - model.build_inducing_kl())
    loss.backward()
    return loss.detach(), {k: v.detach().clone() for k, v in product_grads(model, "shooting").items()}


def main():
    from gaussian_process_odes_b200 import distributed
    rank, world, local = distributed.init_from_env()
    assert world >= 2 and dist.get_backend() == "nccl"
    worst = 0.0

    # ---- A: row sharding -------------------------------------------------------------------------------------------
    kw = dict(D=2, M=16, S=256, N=1, T=25, S_mc=5)
    p, ys, ts, draws, proj = O.make_problem(seed=21, **kw)
    l_full, g_full = unsharded(p, ys, ts, draws, proj, kw)
    model = build_product_model("shooting", p, ys, kw['S'], "rk4")
    distributed.enable_row_sharding(model)
    assert model.row_shard == (rank, world)
    with injected_draws(draws, mvn_order=("eps_x0", "eps_states")):
        loss = distributed.row_sharded_shooting_loss(model, ys.cuda(), ts.cuda(), kw['S_mc'])
    loss.backward()
    n = distributed.allreduce_all_grads(model)
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    assert relerr(tot, l_full) <= 1e-6, ("row-sharded loss", float(tot), float(l_full))
    for k, v in product_grads(model, "shooting").items():
        e = relerr(v, g_full[k])
        worst = max(worst, e)
        assert e <= 2e-5, ("row-sharded grad", k, e)
    if rank == 0:
        print("A row sharding ok: %d gradient floats all-reduced, worst gradient deviation %.2e" % (n, worst))

    # ---- A2: time sharding -----------------------------------------------------------------------------------------
    for kw2, seed in ((dict(D=2, M=16, S=256, N=1, T=25, S_mc=5), 21), (dict(D=3, M=24, S=64, N=2, T=41, S_mc=2, D_obs=7), 23)):
        p, ys, ts, draws, proj = O.make_problem(seed=seed, **kw2)
        l_full, g_full = unsharded(p, ys, ts, draws, proj, kw2)
        model = build_product_model("shooting", p, ys, kw2['S'], "rk4", proj=None if proj is None else proj.components)
        distributed.enable_time_sharding(model)
        assert model.time_shard == (rank, world) and model.row_shard is None
        with injected_draws(draws, mvn_order=("eps_x0", "eps_states")):
            loss = distributed.time_sharded_shooting_loss(model, ys.cuda(), ts.cuda(), kw2['S_mc'])
        loss.backward()
        distributed.allreduce_all_grads(model)
        tot = loss.detach().clone()
        dist.all_reduce(tot)
        assert relerr(tot, l_full) <= 1e-6, ("time-sharded loss", float(tot), float(l_full))
        worst = 0.0
        for k, v in product_grads(model, "shooting").items():
            e = relerr(v, g_full[k])
            worst = max(worst, e)
            assert e <= 2e-5, ("time-sharded grad", k, e)
        if rank == 0:
            print("A2 time sharding ok (N=%d, T=%d): worst gradient deviation %.2e" % (kw2['N'], kw2['T'], worst))

    # ---- B: sequence sharding ----------------------------------------------------------------------------------------
    kw = dict(D=5, M=100, S=256, N=2 * world, T=40, S_mc=3, D_obs=50, dt=0.01, ell0=1.25)
    p, ys, ts, draws, proj = O.make_problem(seed=22, **kw)
    l_full, g_full = unsharded(p, ys, ts, draws, proj, kw)
    lo, hi = distributed.shard_range(kw['N'], rank, world)
    local_keys = ("x0_mean", "x0_lchol_packed", "state_mean", "state_lchol_packed")
    p_loc = {k: (v[lo:hi].clone() if k in local_keys else v) for k, v in p.items()}
    d_loc = dict(draws)
    d_loc['eps_x0'], d_loc['eps_states'] = draws['eps_x0'][:, lo:hi], draws['eps_states'][:, lo:hi]
    model = build_product_model("shooting", p_loc, ys[lo:hi], kw['S'], "rk4", proj=proj.components)
    model.num_observations = kw['N'] * kw['T'] * kw['D_obs']
    with injected_draws(d_loc, mvn_order=("eps_x0", "eps_states")):
        loss = distributed.sharded_shooting_loss(model, ys[lo:hi].cuda(), ts.cuda(), kw['S_mc'], kw['N'], world)
    loss.backward()
    distributed.allreduce_shared_grads(model)
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    assert relerr(tot, l_full) <= 1e-6, ("sequence-sharded loss", float(tot), float(l_full))
    worst = 0.0
    for k, v in product_grads(model, "shooting").items():
        ref = g_full[k][lo:hi] if k in local_keys else g_full[k]
        e = float((v - ref).abs().max() / (g_full[k].abs().max() + 1e-300))
        worst = max(worst, e)
        assert e <= 2e-5, ("sequence-sharded grad", k, e)
    if rank == 0:
        print("B sequence sharding ok: worst gradient deviation %.2e" % worst)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIRANK_OK")


if __name__ == "__main__":
    main()
