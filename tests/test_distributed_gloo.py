"""CPU, world_size 2, gloo: the N>1 path's host logic -- sequence sharding, the rank-local loss whose sum is the
global ELBO, and the single all-reduce of the shared-parameter gradient. The per-shard compute is done by the oracle
port here (the CUDA kernels need a GPU); the gradients are planted into the PRODUCT model's parameters and reduced by
the product's own ``allreduce_shared_grads``."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(D=2, M=8, S=32, N=4, T=6, S_mc=3, seed=11)


def _shard_problem(p, ys, draws, lo, hi):
    ps = dict(p)
    for k in ("x0_mean", "x0_lchol_packed", "state_mean", "state_lchol_packed"):
        ps[k] = p[k][lo:hi]
    ds = dict(draws)
    ds["eps_x0"] = draws["eps_x0"][:, lo:hi]
    ds["eps_states"] = draws["eps_states"][:, lo:hi]
    return ps, ys[lo:hi], ds


def _worker(rank, world, port, out):
    for pth in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, pth)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
    import gpode_oracle as O
    from gaussian_process_odes_b200 import builders, distributed
    torch.set_num_threads(1)
    r, w, _ = distributed.init_from_env(backend="gloo")
    p, ys, ts, draws, _ = O.make_problem(dtype=torch.float64, **KW)  # float64: this test is about the decomposition
    N = KW["N"]
    lo, hi = distributed.shard_range(N, r, w)
    ps, ys_l, ds = _shard_problem(p, ys, draws, lo, hi)
    pp = {k: v.detach().clone().requires_grad_(True) for k, v in ps.items()}
    nobs = N * KW["T"] * KW["D"]
    res = O.elbo_shooting(pp, ys_l, ts, ds, method="rk4", num_observations=nobs)
    loss_r = distributed.combine_shard_terms(res["observ_loglik"], res["constraint_loglik"], res["state_entropy"],
                                             res["init_state_kl"], res["inducing_kl"], hi - lo, N, w)
    loss_r.backward()
    # plant the shard gradients into a product model and reduce with the product's own routine
    model = builders.build_gpode_shooting(hi - lo, KW["T"], KW["D"], num_inducing=KW["M"], num_features=KW["S"],
                                          solver="rk4")
    gp = model.flow.odefunc.diffeq
    pairs = {"inducing_loc": gp.inducing_loc.optvar, "Um": gp.Um.optvar, "Us_sqrt_packed": gp.Us_sqrt.optvar,
             "unconstrained_lengthscales": gp.kern.unconstrained_lengthscales,
             "unconstrained_variance": gp.kern.unconstrained_variance,
             "lik_unconstrained_variance": model.likelihood.unconstrained_variance}
    for k, prm in pairs.items():
        prm.grad = pp[k].grad.clone().to(prm.dtype)
    model.state_distribution.param_mean.optvar.grad = pp["state_mean"].grad.clone().float()
    shared = distributed.shared_parameters(model)
    assert all(id(model.state_distribution.param_mean.optvar) != id(q) for q in shared)
    n = distributed.allreduce_shared_grads(model)
    total = torch.tensor([float(loss_r.detach())], dtype=torch.float64)
    torch.distributed.all_reduce(total)
    if r == 0:
        torch.save(dict(n=n, loss=total, grads={k: prm.grad.clone() for k, prm in pairs.items()},
                        state_mean_grad=pp["state_mean"].grad.clone()), out)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_elbo_and_gradient_allreduce(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gpode_oracle as O
    out = str(tmp_path / "rank0.pt")
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    p, ys, ts, draws, _ = O.make_problem(dtype=torch.float64, **KW)
    pp = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    ref = O.elbo_shooting(pp, ys, ts, draws, method="rk4")
    ref["loss"].backward()
    assert abs(float(got["loss"]) - float(ref["loss"].detach())) <= 1e-9 * abs(float(ref["loss"].detach()))
    assert got["n"] > 0
    for k, g in got["grads"].items():
        err = float((g - pp[k].grad).abs().max() / (pp[k].grad.abs().max() + 1e-30))
        assert err <= 1e-6, (k, err)
    # per-sequence state gradients stay local and equal the global gradient's block
    lo, hi = 0, 2
    err = float((got["state_mean_grad"] - pp["state_mean"].grad[lo:hi]).abs().max()
                / pp["state_mean"].grad.abs().max())
    assert err <= 1e-9


def _pred_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
    from gaussian_process_odes_b200 import distributed
    torch.set_num_threads(1)
    r, w, _ = distributed.init_from_env(backend="gloo")
    S, N, T, D = 7, 2, 5, 3  # 7 draws over 2 ranks: blocks of 4 and 3

    def fake_predict(model, ts, eval_sample_size):  # stands in for the CUDA n_sets path: draw q -> constant q
        lo, _ = distributed.shard_range(S, r, w)
        ids = torch.arange(lo, lo + eval_sample_size, dtype=torch.float32)
        return ids.view(-1, 1, 1, 1).expand(eval_sample_size, N, T, D).contiguous()

    full = distributed.sharded_predictions(None, None, S, predict_fn=fake_predict)
    if r == 1:
        torch.save(full, out)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_predictions_gathers_draws_in_rank_order(tmp_path):
    out = str(tmp_path / "rank1.pt")
    port = 29950 + os.getpid() % 40
    mp.spawn(_pred_worker, args=(2, port, out), nprocs=2, join=True)
    full = torch.load(out)
    assert full.shape == (7, 2, 5, 3)
    assert torch.equal(full[:, 0, 0, 0], torch.arange(7, dtype=torch.float32))
