"""CPU: host-side logic of the reference-API mirror (no kernel launches)."""
import numpy as np
import pytest
import torch

import gpode_oracle as O


def test_tril_scatter_matches_reference_loop_semantics():
    from gaussian_process_odes_b200.misc import transforms
    rng = np.random.default_rng(0)
    n, k = 4, 3
    x = torch.tensor(rng.normal(size=(k, n * (n + 1) // 2)), dtype=torch.float32)
    tr = transforms.LowerTriangular(n, k)
    out = tr.forward_tensor(x)
    r, c = np.tril_indices(n)
    ref = torch.zeros(k, n, n)
    for i in range(k):  # the reference's loop (src/misc/transforms.py:70-76)
        ref[i, r, c] = x[i]
    assert torch.equal(out, ref)
    assert torch.equal(tr.backward_tensor(out), x)
    assert np.array_equal(tr.forward(x.numpy()), ref.numpy())
    assert np.array_equal(tr.backward(ref.numpy()), x.numpy())
    st = transforms.StackedLowerTriangular(n, 2, 5)
    y = torch.tensor(rng.normal(size=(2, 5, n * (n + 1) // 2)), dtype=torch.float32)
    out2 = st.forward_tensor(y)
    for i in range(2):
        for j in range(5):  # src/misc/transforms.py:105-112
            m = torch.zeros(n, n)
            m[r, c] = y[i, j]
            assert torch.equal(out2[i, j], m)
    assert torch.equal(st.backward_tensor(out2), y)
    # gradient flows through the scatter
    y2 = y.clone().requires_grad_(True)
    st.forward_tensor(y2).sum().backward()
    assert torch.equal(y2.grad, torch.ones_like(y))


def test_step_grids_match_reference_formulas():
    from gaussian_process_odes_b200.misc.torch_utils import compute_ts_dense, insert_zero_t0
    ts = torch.linspace(0, 7, 25)
    z = insert_zero_t0(ts)
    assert torch.equal(z, torch.cat([torch.tensor([0.0]), ts + ts[1] - ts[0]]))  # src/misc/torch_utils.py:36-38
    for k in (2, 4):
        d = compute_ts_dense(z, k)
        ref = torch.cat([torch.linspace(t1, t2, k)[:-1] for (t1, t2) in zip(z[:-1], z[1:])] + [z[-1:]])
        assert torch.equal(d, ref)  # src/misc/torch_utils.py:41-48
        assert torch.equal(d, O.compute_ts_dense(O.insert_zero_t0(ts), k))
        assert d.shape[0] == (len(z) - 1) * (k - 1) + 1
        assert torch.equal(d[::k - 1], z)
    assert compute_ts_dense(ts, 1) is ts
    # cache must notice in-place edits
    ts2 = ts.clone()
    a = insert_zero_t0(ts2)
    ts2.mul_(2.0)
    b = insert_zero_t0(ts2)
    assert not torch.equal(a, b)


def test_state_dict_names_match_reference_modules():
    from gaussian_process_odes_b200 import builders
    m = builders.build_gpode(1, 25, 2)
    assert set(m.state_dict().keys()) == {
        'flow.odefunc._num_evals', 'flow.odefunc.diffeq.kern.unconstrained_lengthscales',
        'flow.odefunc.diffeq.kern.unconstrained_variance', 'flow.odefunc.diffeq.inducing_loc.optvar',
        'flow.odefunc.diffeq.Um.optvar', 'flow.odefunc.diffeq.Us_sqrt.optvar', 'x0_distribution.param_mean.optvar',
        'x0_distribution.param_lchol.optvar', 'likelihood.unconstrained_variance'}
    sd = m.state_dict()
    assert sd['flow.odefunc.diffeq.Us_sqrt.optvar'].shape == (2, 16 * 17 // 2)
    assert sd['flow.odefunc.diffeq.kern.unconstrained_lengthscales'].shape == (2, 2)
    gp = m.flow.odefunc.diffeq
    assert abs(float(gp.kern.lengthscales[0, 0]) - 1.3) < 1e-5 and abs(float(gp.kern.variance[0]) - 0.5) < 1e-6
    assert torch.allclose(gp.Us_sqrt().cpu(), torch.eye(16).expand(2, 16, 16) * 1e-3, atol=1e-9)


def test_side_terms_match_oracle_on_cpu():
    """states / likelihood / constraint are plain torch: check them against the oracle port."""
    from gaussian_process_odes_b200.core import constraints, likelihoods, states
    if torch.cuda.is_available():
        pytest.skip("module parameters live on the GPU here; the gpu tests cover these terms through the ELBO")
    p, ys, ts, draws, _ = O.make_problem(D=3, M=8, S=16, N=2, T=6, S_mc=4, seed=3)
    sd = states.StateSequenceVariationalFactorizedGaussian(2, 5, 3)
    with torch.no_grad():
        sd.param_mean.optvar.copy_(p['state_mean'])
        sd.param_lchol.optvar.copy_(p['state_lchol_packed'] + 0.01 * torch.randn_like(p['state_lchol_packed']))
        sd.x0.param_mean.optvar.copy_(p['x0_mean'])
        sd.x0.param_lchol.optvar.copy_(p['x0_lchol_packed'])
    q = [draws['eps_x0'], draws['eps_states']]
    saved = states._standard_normal
    states._standard_normal = lambda shape, dtype, device: q.pop(0)
    try:
        ss = sd.sample(4)
    finally:
        states._standard_normal = saved
    d0 = O.mvn_from_lchol(p['x0_mean'], O.tril_from_packed(p['x0_lchol_packed'], 3))
    ds = O.mvn_from_lchol(sd.mean().detach(), sd.lchol().detach())
    ref = torch.cat([O.mvn_rsample(d0, draws['eps_x0']).unsqueeze(2), O.mvn_rsample(ds, draws['eps_states'])], 2)
    assert torch.allclose(ss, ref, atol=1e-6)
    assert torch.allclose(sd.entropy(), ds.entropy(), atol=1e-5)
    assert torch.allclose(sd.x0.kl(), O.x0_kl(p['x0_mean'], O.tril_from_packed(p['x0_lchol_packed'], 3)), atol=1e-6)
    x = torch.randn(2, 5, 3)
    assert torch.allclose(sd.log_prob(x), ds.log_prob(x), atol=1e-4)
    lik = likelihoods.Gaussian(ndim=3)
    f, y = torch.randn(4, 2, 6, 3), torch.randn(1, 2, 6, 3)
    assert torch.allclose(lik.log_prob(f, y), O.gauss_loglik(f, y, lik.variance), atol=1e-6)
    cg = constraints.Gaussian(d=1, scale=1e-3, requires_grad=False)
    assert torch.allclose(cg.log_prob(f, y), O.normal_logprob(y, f, cg.scale), rtol=1e-5)
    cl = constraints.Laplace(d=1, scale=0.5)
    assert torch.allclose(cl.log_prob(f, y), torch.distributions.Laplace(f, cl.scale).log_prob(y), atol=1e-6)


def test_no_cpu_fallback():
    """Without a CUDA device the hot path must raise, never compute on the CPU."""
    from gaussian_process_odes_b200 import builders, _lib
    if torch.cuda.is_available():
        pytest.skip("this check is for CPU-only machines")
    m = builders.build_gpode_shooting(1, 5, 2, num_inducing=4, num_features=8, solver='rk4')
    with pytest.raises(_lib.GpodeError):
        m.flow.odefunc.diffeq.build_cache()
    with pytest.raises(_lib.GpodeError):
        m.build_lowerbound_terms(torch.zeros(1, 5, 2), torch.linspace(0, 1, 5), num_samples=2)


def test_odeint_rejects_foreign_functions_and_solvers():
    from gaussian_process_odes_b200 import _lib
    from gaussian_process_odes_b200.odeint import odeint
    with pytest.raises(_lib.GpodeError):
        odeint(torch.nn.Linear(2, 2), torch.zeros(1, 2), torch.tensor([0.0, 1.0]), method='rk4')
    with pytest.raises(_lib.GpodeError):
        odeint(torch.nn.Linear(2, 2), torch.zeros(1, 2), torch.tensor([0.0, 1.0]), method='adams')


def test_shard_ranges_cover_everything():
    from gaussian_process_odes_b200.distributed import shard_range
    for n in (1, 7, 16, 125):
        for world in (1, 2, 3, 8):
            got = [shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
