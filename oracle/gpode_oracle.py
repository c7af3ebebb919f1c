"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle port") of GPODE's hot path.

Every function cites the reference file:line it follows (paths relative to the reference repo root).
It is written with plain torch CPU ops in the SAME operation order as the reference so that the float32
result is (near) bit-identical to the reference's own PyTorch path, and it accepts ``dtype=torch.float64``
to serve as the arbiter for float32 round-off (SURVEY.md section 7 "Tolerance vs fp32 noise").

Pinning: ``oracle/pin_against_reference.py`` runs the UNMODIFIED reference modules (imported from
``/root/reference`` with the ``torchdiffeq`` shim and injected random draws, see ``oracle/reference_harness.py``)
and checks this restatement against them; the same script writes the committed fixtures under ``tests/golden/``.
The integrator arithmetic itself (torchdiffeq 0.2.0) is a third-party dependency that is absent offline, so that
part is PARITY UNPINNED (see ``oracle/torchdiffeq_shim``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
Nothing under ``gaussian_process_odes_b200/`` may import it: the product path has no CPU fallback.
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "torchdiffeq_shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)
import torchdiffeq as _tde  # noqa: E402  (the restated torchdiffeq 0.2.0)

JITTER = 1e-5  # src/core/dsvgp.py:8, src/core/states.py:11


# ----------------------------------------------------------------------------------------------------------------
# parameter plumbing
# ----------------------------------------------------------------------------------------------------------------
def softplus(x):
    """src/misc/constraint_utils.py:5-7"""
    return F.softplus(x) + 1e-12


def invsoftplus(x):
    """src/misc/constraint_utils.py:10-13"""
    x = torch.as_tensor(x)
    xs = torch.max(x - 1e-12, torch.tensor(torch.finfo(x.dtype).eps).to(x))
    return xs + torch.log(-torch.expm1(-xs))


def tril_from_packed(packed, n):
    """src/misc/transforms.py:70-76 and :105-112 -- row-major ``np.tril_indices`` scatter of the last axis
    ``(..., n(n+1)/2) -> (..., n, n)`` (vectorised; the reference loops in Python, the result is identical)."""
    r, c = np.tril_indices(n, 0)
    out = packed.new_zeros(packed.shape[:-1] + (n, n))
    out[..., torch.as_tensor(r), torch.as_tensor(c)] = packed
    return out


def packed_from_tril(mat):
    """src/misc/transforms.py:66-68,101-103"""
    n = mat.shape[-1]
    r, c = np.tril_indices(n)
    return mat[..., torch.as_tensor(r), torch.as_tensor(c)]


def insert_zero_t0(ts):
    """src/misc/torch_utils.py:36-38"""
    return torch.cat([torch.zeros(1, dtype=ts.dtype), ts + ts[1] - ts[0]])


def compute_ts_dense(ts, ts_dense_scale):
    """src/misc/torch_utils.py:41-48"""
    if ts_dense_scale > 1:
        return torch.cat([torch.linspace(float(t1), float(t2), ts_dense_scale, dtype=ts.dtype)[:-1]
                          for (t1, t2) in zip(ts[:-1], ts[1:])] + [ts[-1:]])
    return ts


# ----------------------------------------------------------------------------------------------------------------
# RBF kernel (src/core/kernels.py)
# ----------------------------------------------------------------------------------------------------------------
def square_dist_dimwise(X, X2, ell):
    """src/core/kernels.py:53-68 -- expanded form, ``ell`` is (D_out, D_in); returns (D_out, N, M)."""
    X = X.unsqueeze(0) / ell.unsqueeze(1)
    Xs = torch.sum(torch.pow(X, 2), dim=2)
    if X2 is None:
        return -2 * torch.einsum('dnk, dmk -> dnm', X, X) + Xs.unsqueeze(-1) + Xs.unsqueeze(1)
    X2 = X2.unsqueeze(0) / ell.unsqueeze(1)
    X2s = torch.sum(torch.pow(X2, 2), dim=2)
    return -2 * torch.einsum('dnk, dmk -> dnm', X, X2) + Xs.unsqueeze(-1) + X2s.unsqueeze(1)


def rbf_K(X, X2, ell, var):
    """src/core/kernels.py:87-99 (dimwise=True branch)."""
    return var[:, None, None] * torch.exp(-0.5 * square_dist_dimwise(X, X2, ell))


# ----------------------------------------------------------------------------------------------------------------
# DSVGP layer (src/core/dsvgp.py)
# ----------------------------------------------------------------------------------------------------------------
def rff_forward(x, omega, phase, w, var):
    """src/core/dsvgp.py:124-137 (dimwise): omega (D_in,S,D_out), phase (1,S,D_out), w (S,D_out)."""
    S = omega.shape[1]
    xo = torch.einsum('nd,dfk->nfk', x, omega)
    phi_ = torch.cos(xo + phase)
    phi = phi_ * torch.sqrt(var / S)
    return torch.einsum('nfk,fk->nk', phi, w)


def sample_inducing(Um, Us_sqrt, eps_u):
    """src/core/dsvgp.py:78-90 (q_diag=False): Us_sqrt (D,M,M), eps_u (M,D) -> (M,D)."""
    return torch.einsum('dnm, md->nd', Us_sqrt, eps_u) + Um


def build_cache(Z, Um, Us_sqrt, ell, var, w, eps_omega, phase_u, eps_u):
    """src/core/dsvgp.py:92-122 with the four random draws passed in:
    ``w`` ~ N(0,1) (S,D) [:100], ``eps_omega`` ~ N(0,1) (D,S,D) [:101, kernels.py:108-112],
    ``phase_u`` ~ U(0,1) (1,S,D) [:102-103], ``eps_u`` ~ N(0,1) (M,D) [:83]."""
    M = Z.shape[0]
    omega = eps_omega / ell.T.unsqueeze(1)  # kernels.py:110-112
    phase = phase_u * 2 * np.pi
    u = sample_inducing(Um, Us_sqrt, eps_u)
    Ku = rbf_K(Z, None, ell, var)
    Lu = torch.linalg.cholesky(Ku + torch.eye(M, dtype=Z.dtype) * JITTER)
    u_prior = rff_forward(Z, omega, phase, w, var)
    nu = torch.linalg.solve_triangular(Lu, u_prior.T.unsqueeze(2), upper=False)
    nu = torch.linalg.solve_triangular(Lu.permute(0, 2, 1), u.T.unsqueeze(2) - nu, upper=True)
    return dict(rff_weights=w, rff_omega=omega, rff_phase=phase, nu=nu, Lu=Lu, u=u)


def vf_forward(x, Z, ell, var, cache):
    """src/core/dsvgp.py:172-197 -- the GP vector field f(x) (dimwise)."""
    f_prior = rff_forward(x, cache['rff_omega'], cache['rff_phase'], cache['rff_weights'], var)
    Kuf = rbf_K(Z, x, ell, var)
    f_update = torch.einsum('dm, dmn -> nd', cache['nu'].squeeze(2), Kuf)
    return f_prior + f_update


def vf_closed_form(x, Z, ell, var, omega, phase, w, nu):
    """The closed form of SURVEY.md section 8(a) row A1, with the squared distance in DIRECT form
    (no cancellation). Used in float64 as the round-off arbiter."""
    S = omega.shape[1]
    a = w * torch.sqrt(var / S)  # (S,K)
    theta = torch.einsum('nj,jsk->nsk', x, omega) + phase.reshape(1, S, -1)
    f_rff = (a.unsqueeze(0) * torch.cos(theta)).sum(1)
    d = (x[:, None, None, :] - Z[None, None, :, :]) / ell[None, :, None, :]  # (N,K,M,J)
    Kxz = torch.exp(-0.5 * (d * d).sum(-1))  # (N,K,M)
    f_upd = (Kxz * (var[:, None] * nu.reshape(ell.shape[0], -1))[None]).sum(-1)
    return f_rff + f_upd


def kl_whitened(Um, Us_sqrt):
    """src/core/dsvgp.py:199-230 (q_diag=False)."""
    M = Um.shape[0]
    Lq = torch.tril(Us_sqrt)
    Lq_diag = torch.diagonal(Lq, dim1=1, dim2=2).t()
    mahalanobis = torch.pow(Um, 2).sum(dim=0, keepdim=True)
    logdet_qcov = torch.log(torch.pow(Lq_diag, 2)).sum(dim=0, keepdim=True)
    trace = torch.pow(Lq, 2).sum(dim=(1, 2)).unsqueeze(0)
    twoKL = -logdet_qcov + mahalanobis + trace - float(M)
    return 0.5 * twoKL.sum()


# ----------------------------------------------------------------------------------------------------------------
# integrator (src/core/flow.py:60-90 calling torchdiffeq 0.2.0)
# ----------------------------------------------------------------------------------------------------------------
def odeint(f, y0, t, method='rk4', rtol=1e-6, atol=1e-6, stats=None):
    """``torchdiffeq.odeint`` semantics (restated in oracle/torchdiffeq_shim); ``f(t, y)``; returns (Tg,B,D)."""
    return _tde.odeint(f, y0, t, rtol=rtol, atol=atol, method=method, _stats=stats)


def flow_forward(x0, ts, gp, cache, method='rk4', rtol=1e-6, atol=1e-6, stats=None):
    """src/core/flow.py:60-90: integrate and return (B,Tg,D)."""
    nfe = [0]

    def f(t, y):
        nfe[0] += 1
        return vf_forward(y, gp['Z'], gp['ell'], gp['var'], cache)

    xs = odeint(f, x0, ts, method=method, rtol=rtol, atol=atol, stats=stats)
    if stats is not None:
        stats['nfe'] = nfe[0]
    return xs.permute(1, 0, 2)


# ----------------------------------------------------------------------------------------------------------------
# ELBO side terms (src/core/states.py, likelihoods.py, constraints.py)
# ----------------------------------------------------------------------------------------------------------------
def mvn_from_lchol(mean, lchol):
    """src/core/states.py:69-74 and :177-182: N(mean, L L^T + 1e-5 I)."""
    cov = lchol @ lchol.transpose(-1, -2)
    cov = cov + torch.eye(cov.shape[-1], dtype=cov.dtype) * JITTER
    return torch.distributions.MultivariateNormal(loc=mean, covariance_matrix=cov)


def mvn_rsample(dist, eps):
    """``MultivariateNormal.rsample`` with the standard-normal draw ``eps`` (S, *batch, D) injected."""
    return dist.loc + (dist._unbroadcasted_scale_tril @ eps.unsqueeze(-1)).squeeze(-1)


def x0_kl(mean, lchol):
    """src/core/states.py:97-114."""
    D = mean.shape[-1]
    Lq = torch.tril(lchol)
    Lq_diag = torch.diagonal(Lq, dim1=1, dim2=2)
    mahalanobis = torch.pow(mean, 2).sum(dim=1, keepdim=True)
    logdet_qcov = torch.log(torch.pow(Lq_diag, 2)).sum(dim=1, keepdim=True)
    trace = torch.pow(Lq, 2).sum(dim=(1, 2)).unsqueeze(1)
    twoKL = -logdet_qcov + mahalanobis + trace - float(D)
    return 0.5 * twoKL.sum()


def gauss_loglik(Fm, Y, variance):
    """src/core/likelihoods.py:27-28."""
    return -0.5 * (np.log(2.0 * np.pi) + torch.log(variance) + torch.pow(Fm - Y, 2) / variance)


def normal_logprob(y, loc, scale):
    """src/core/constraints.py:26-36 (``Normal(loc, scale).log_prob(y)``)."""
    return torch.distributions.Normal(loc=loc, scale=scale).log_prob(y)


# ----------------------------------------------------------------------------------------------------------------
# the two ELBOs
# ----------------------------------------------------------------------------------------------------------------
def gp_params(p):
    """Constrained GP parameters from the module-state layout of SURVEY.md section 8(a) (state_dict names).
    ``Us_sqrt_diag_unconstrained`` (M,D) instead of ``Us_sqrt_packed`` selects q_diag=True (src/core/dsvgp.py:69-72):
    the diagonal scale is embedded as D diagonal (M,M) matrices, for which sample_inducing and kl give the q_diag
    formulas of src/core/dsvgp.py:84-85,207-223."""
    M, D = p['inducing_loc'].shape
    if 'Us_sqrt_diag_unconstrained' in p:
        Us = torch.diag_embed(softplus(p['Us_sqrt_diag_unconstrained']).t())  # (D,M,M)
    else:
        Us = tril_from_packed(p['Us_sqrt_packed'], M)
    return dict(Z=p['inducing_loc'], Um=p['Um'], Us_sqrt=Us,
                ell=softplus(p['unconstrained_lengthscales']), var=softplus(p['unconstrained_variance']))


def elbo_gpode(p, ys, ts, draws, ts_dense_scale=4, method='rk4', project=None, num_observations=None, stats=None):
    """src/gpode/model_builder.py:46-57 + src/gpode/models.py:32-66.
    ``p``: dict of UNCONSTRAINED tensors (state_dict layout); ``draws``: the injected random numbers."""
    N, T, Dobs = ys.shape
    gp = gp_params(p)
    D = gp['Z'].shape[1]
    nobs = num_observations if num_observations is not None else N * T * Dobs
    # the step grid is always formed in float32 (the reference's dtype) so a float64 arbiter run integrates
    # over exactly the same grid values
    tsd = compute_ts_dense(insert_zero_t0(ts.float()), ts_dense_scale).to(ys.dtype)
    x0_mean, x0_lchol = p['x0_mean'], tril_from_packed(p['x0_lchol_packed'], D)
    x0 = mvn_rsample(mvn_from_lchol(x0_mean, x0_lchol), draws['eps_x0'])[0]
    kl0 = x0_kl(x0_mean, x0_lchol)
    cache = build_cache(gp['Z'], gp['Um'], gp['Us_sqrt'], gp['ell'], gp['var'],
                        draws['w'], draws['eps_omega'], draws['phase_u'], draws['eps_u'])
    xs = flow_forward(x0, tsd, gp, cache, method=method, stats=stats)[:, ::ts_dense_scale - 1, :][:, 1:]
    pred = project(xs) if project is not None else xs
    loglik = gauss_loglik(pred, ys, softplus(p['lik_unconstrained_variance']))
    kl_u = kl_whitened(gp['Um'], gp['Us_sqrt']) / nobs
    ll, k0 = loglik.mean(), kl0.mean() / nobs
    loss = -(ll - k0 - kl_u)
    return dict(loss=loss, observ_loglik=ll, init_state_kl=k0, inducing_kl=kl_u, xs=xs, x0=x0, cache=cache)


def elbo_shooting(p, ys, ts, draws, method='rk4', project=None, num_observations=None, stats=None):
    """src/gpode_shooting/model_builder.py:59-72 + src/gpode_shooting/models.py:108-146 (Gaussian constraint)."""
    N, T, Dobs = ys.shape
    gp = gp_params(p)
    D = gp['Z'].shape[1]
    nobs = num_observations if num_observations is not None else N * T * Dobs
    x0_mean, x0_lchol = p['x0_mean'], tril_from_packed(p['x0_lchol_packed'], D)
    s_mean, s_lchol = p['state_mean'], tril_from_packed(p['state_lchol_packed'], D)
    d0, ds = mvn_from_lchol(x0_mean, x0_lchol), mvn_from_lchol(s_mean, s_lchol)
    ss = torch.cat([mvn_rsample(d0, draws['eps_x0']).unsqueeze(2), mvn_rsample(ds, draws['eps_states'])], 2)
    S = ss.shape[0]
    cache = build_cache(gp['Z'], gp['Um'], gp['Us_sqrt'], gp['ell'], gp['var'],
                        draws['w'], draws['eps_omega'], draws['phase_u'], draws['eps_u'])
    pred = flow_forward(ss.reshape(-1, D), ts[:2], gp, cache, method=method, stats=stats)[:, -1].reshape(S, N, T, D)
    if project is not None:
        proj = torch.stack([project(_F) for _F in pred])
    else:
        proj = pred
    loglik = gauss_loglik(proj, ys.unsqueeze(0), softplus(p['lik_unconstrained_variance']))
    entropy = ds.entropy()
    cons = normal_logprob(ss[:, :, 1:, :], pred[:, :, :-1, :], softplus(p['constraint_unconstrained_scale'])).sum(3)
    kl0 = x0_kl(x0_mean, x0_lchol)
    ll = loglik.mean()
    c = cons.mean(0).sum() / nobs
    e = entropy.sum() / nobs
    k0 = kl0 / nobs
    kl_u = kl_whitened(gp['Um'], gp['Us_sqrt']) / nobs
    loss = -(ll + c + e - k0 - kl_u)
    return dict(loss=loss, observ_loglik=ll, constraint_loglik=c, state_entropy=e, init_state_kl=k0,
                inducing_kl=kl_u, pred=pred, ss=ss, cache=cache)


# ----------------------------------------------------------------------------------------------------------------
# synthetic problems of the BASELINE configs; shared by tests, smoke and bench
# ----------------------------------------------------------------------------------------------------------------
def make_problem(D, M, S, N, T, seed=121, S_mc=1, D_obs=None, dt=None, t_end=7.0, ell0=1.3, var0=0.5,
                 dtype=torch.float32):
    """Unconstrained parameters in the reference's init style (src/core/dsvgp.py:66-76, kernels.py:41-43,
    states.py:57-63,159-166, likelihoods.py:15) + synthetic data + one set of injected draws."""
    rng = np.random.default_rng(seed)
    t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
    Dobs = D if D_obs is None else D_obs
    p = dict(
        inducing_loc=t(rng.normal(size=(M, D)) * 1.5),
        Um=t(rng.normal(size=(M, D)) * 1e-1),
        Us_sqrt_packed=packed_from_tril(t(np.stack([np.eye(M)] * D) * 1e-3 + np.tril(rng.normal(size=(D, M, M))) * 1e-4)),
        unconstrained_lengthscales=invsoftplus(t(ell0 * (1 + 0.1 * rng.normal(size=(D, D))))),
        unconstrained_variance=invsoftplus(t(var0 * (1 + 0.1 * rng.normal(size=(D,))))),
        x0_mean=t(rng.normal(size=(N, D))),
        x0_lchol_packed=packed_from_tril(t(np.stack([np.eye(D)] * N) * 1e-1)),
        state_mean=t(rng.normal(size=(N, T - 1, D))),
        state_lchol_packed=packed_from_tril(t(np.stack([np.stack([np.eye(D)] * (T - 1))] * N) * 1e-1)),
        lik_unconstrained_variance=invsoftplus(t(np.full((Dobs,), 0.25))),
        constraint_unconstrained_scale=invsoftplus(t(np.full((1,), 1e-3))),
    )
    ys = t(rng.normal(size=(N, T, Dobs)) * 1.5)
    ts = t(np.arange(T) * dt) if dt is not None else t(np.linspace(0.0, t_end, T))
    draws = make_draws(D, M, S, N, T, S_mc, rng, dtype)
    proj = None
    if D_obs is not None:
        q, _ = np.linalg.qr(rng.normal(size=(D_obs, D)))
        comp = t(q.T)  # (D, D_obs) orthonormal rows, the fixed inverse-PCA decoder (src/misc/mocap_utils.py:29)
        proj = lambda x: torch.einsum('ntl,ld->ntd', x, comp.to(x.dtype))
        proj.components = comp
    return p, ys, ts, draws, proj


def make_draws(D, M, S, N, T, S_mc, rng, dtype=torch.float32):
    t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
    return dict(w=t(rng.normal(size=(S, D))), eps_omega=t(rng.normal(size=(D, S, D))),
                phase_u=t(rng.uniform(size=(1, S, D))), eps_u=t(rng.normal(size=(M, D))),
                eps_x0=t(rng.normal(size=(S_mc, N, D))), eps_states=t(rng.normal(size=(S_mc, N, T - 1, D))))


def cast(tree, dtype):
    if isinstance(tree, dict):
        return {k: cast(v, dtype) for k, v in tree.items()}
    if torch.is_tensor(tree) and tree.is_floating_point():
        return tree.detach().to(dtype)
    return tree
