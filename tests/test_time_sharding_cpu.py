"""CPU: the host-side pieces of TIME sharding (distributed.enable_time_sharding) -- the slice of the sampled states each
process computes (own time indices + one halo state), its entropy share and where the halo's gradient lands. The
integrator needs a GPU (tests/mp_nccl_worker.py, case A2, runs the sharded ELBO on two B200s); here the product's own
state-distribution classes run on the CPU and the slices are compared with the unsharded sample / entropy
(reference src/core/states.py:144-207, src/gpode_shooting/models.py:119-135)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _state_distribution(N, T, D, seed):
    from gaussian_process_odes_b200.core.states import StateSequenceVariationalFactorizedGaussian
    torch.manual_seed(seed)
    sd = StateSequenceVariationalFactorizedGaussian(N, T, D)
    with torch.no_grad():   # non-trivial factors: the default scale is a multiple of the identity
        for prm in sd.parameters():
            prm.add_(0.1 * torch.randn_like(prm))
    return sd


@pytest.mark.parametrize("N,T,D,S,world", [(3, 6, 2, 4, 2), (1, 25, 2, 5, 8), (2, 5, 3, 2, 4), (1, 3, 2, 2, 8)])
def test_time_slices_reassemble_the_unsharded_sample_and_entropy(N, T, D, S, world):
    from gaussian_process_odes_b200.distributed import shard_range
    sd = _state_distribution(N, T, D, seed=T)
    torch.manual_seed(123)
    full = sd.sample(S)                       # (S, N, T + 1, D): x0 then the T shooting states
    ent_full = sd.entropy()                   # (N, T)
    assert full.shape == (S, N, T + 1, D)
    covered, ent_sum = [], 0.0
    for rank in range(world):
        lo, hi = shard_range(T + 1, rank, world)
        if hi <= lo:
            continue
        halo = hi < T + 1
        torch.manual_seed(123)                # every process seeds alike and draws the full noise in the same order
        loc = sd.sample_time_slice(S, lo, hi + int(halo))
        assert loc.shape == (S, N, hi - lo + int(halo), D)
        assert torch.equal(loc, full[:, :, lo:hi + int(halo)])      # own states and the halo: the SAME numbers
        e = sd.entropy_time_slice(lo, hi)
        a, b = max(lo, 1) - 1, hi - 1
        assert e.shape == (N, max(b - a, 0))
        if b > a:
            assert torch.allclose(e, ent_full[:, a:b], rtol=0, atol=1e-12)
        ent_sum = ent_sum + float(e.detach().double().sum())
        covered += list(range(lo, hi))
    assert covered == list(range(T + 1))      # every time index has exactly one owner
    tot = float(ent_full.detach().double().sum())   # float64 sums: the slices are bitwise the unsharded entries
    assert abs(ent_sum - tot) <= 1e-12 * max(1.0, abs(tot))


def test_halo_gradient_lands_on_the_next_slice_owner_rows():
    """A term that depends on the halo state only (the constraint's neighbour) must put its gradient on the parameter
    rows of the NEXT slice and nowhere else: the all-reduce of all gradients then delivers it to its owner."""
    from gaussian_process_odes_b200.distributed import shard_range
    N, T, D, S, world = 2, 7, 2, 3, 2
    sd = _state_distribution(N, T, D, seed=5)
    lo, hi = shard_range(T + 1, 0, world)
    assert 0 < hi < T + 1
    torch.manual_seed(9)
    loc = sd.sample_time_slice(S, lo, hi + 1)
    loc[:, :, -1].pow(2).sum().backward()     # the halo = time index hi = shooting state hi - 1
    g_mean = sd.param_mean.optvar.grad
    g_chol = sd.param_lchol.optvar.grad
    rows = torch.zeros(T, dtype=torch.bool)
    rows[hi - 1] = True
    assert g_mean[:, rows].abs().sum() > 0 and g_chol[:, rows].abs().sum() > 0
    assert g_mean[:, ~rows].abs().sum() == 0 and g_chol[:, ~rows].abs().sum() == 0
    assert sd.x0.param_mean.optvar.grad is None or sd.x0.param_mean.optvar.grad.abs().sum() == 0


def test_enable_time_sharding_sets_and_clears_the_shard():
    from gaussian_process_odes_b200 import builders, distributed
    model = builders.build_gpode_shooting(1, 6, 2, num_inducing=4, num_features=8, solver="rk4")
    distributed.enable_row_sharding(model, 1, 4)
    distributed.enable_time_sharding(model, 1, 4)
    assert model.time_shard == (1, 4) and model.row_shard is None
    distributed.enable_time_sharding(model, 0, 1)
    assert model.time_shard is None
