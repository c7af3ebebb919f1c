"""Data-driven initialisation (mirror of reference ``src/gpode/model_initialization.py``,
``src/gpode_shooting/model_initialization.py`` and the two ``mocap_initialization.py``).

One-off work before training, so the GP regression is plain ``torch.linalg`` on the model's device; the part that IS
the hot path -- 20-50 backward-in-time integrations, one new GP function draw each (``model_initialization.py:70-75``)
-- goes through the batched ``forward_sets`` path: one whitening, one pack and one integrator launch for all draws.
Works for both model families (``SequenceModel`` has ``x0_distribution``, the shooting model ``state_distribution``).
"""
import numpy as np
import torch

from .misc import constraint_utils


def _device(model):
    return next(model.parameters()).device


def initialize_inducing(model, data_ys, ts_max, data_noise=1e-1):
    """Inducing locations at k-means centres of the observed states, whitened inducing means from a GP regression of
    the empirical gradients ``(y[t+1]-y[t]) * T/ts_max`` on at most 1000 observations (reference
    ``src/gpode/model_initialization.py:6-52``; MoCap variant: ``data_noise=1e0``). ``data_ys``: numpy ``(N,T,D)``.
    The host generator is consumed in the reference's order: the observation subset, then scipy's k-means."""
    from scipy.cluster.vq import kmeans2
    layer = model.flow.odefunc.diffeq
    dev = _device(model)
    data_ys = np.asarray(data_ys)
    D = data_ys.shape[-1]
    f_xt = (data_ys[:, 1:, :] - data_ys[:, :-1, :]).reshape(-1, D) * (data_ys.shape[1] / ts_max)
    xs = data_ys[:, :-1, :].reshape(-1, D)
    with torch.no_grad():
        n_init = int(np.minimum(1000, xs.shape[0]))
        obs_index = np.random.choice(xs.shape[0], n_init, replace=False)
        Z = torch.tensor(kmeans2(xs, k=layer.M, minit='points')[0], dtype=torch.float64, device=dev)
        X = torch.tensor(xs[obs_index], dtype=torch.float64, device=dev)
        F = torch.tensor(f_xt[obs_index], dtype=torch.float64, device=dev)
        ell = layer.kern.lengthscales_dimwise().double()   # (D,D): a shared-lengthscale kernel is its dimwise expansion
        var = layer.kern.variance_dimwise().double()

        def K(A, B):  # (D,|A|,|B|), direct squared-distance form
            d = (A.unsqueeze(0) / ell.unsqueeze(1)).unsqueeze(2) - (B.unsqueeze(0) / ell.unsqueeze(1)).unsqueeze(1)
            return var[:, None, None] * torch.exp(-0.5 * d.pow(2).sum(-1))

        eye = lambda n: torch.eye(n, dtype=torch.float64, device=dev)
        Lxx = torch.linalg.cholesky(K(X, X) + eye(n_init) * data_noise)
        Lzz = torch.linalg.cholesky(K(Z, Z) + eye(layer.M) * 1e-6)
        alpha = torch.cholesky_solve(F.T.unsqueeze(2), Lxx)                       # (D,n,1)
        f_update = torch.einsum('dnm,dn->md', K(X, Z), alpha.squeeze(2))          # (M,D)
        u = torch.linalg.solve_triangular(Lzz, f_update.T.unsqueeze(2), upper=False).squeeze(2).T  # whitened (M,D)
        layer.inducing_loc.optvar.data = Z.to(layer.inducing_loc.optvar)
        layer.Um.optvar.data = u.to(layer.Um.optvar)
    return model


def initial_state_from_data(model, y_first, data_ts, num_samples=20, rng="numpy"):
    """Mean over ``num_samples`` GP draws of the state one sampling interval BEFORE the first observation: the ODE
    solved backward in time from ``y_first (N,D)`` over ``[ts[1], ts[0]]`` (reference
    ``model_initialization.py:66-75``). All draws in one launch."""
    dev = _device(model)
    with torch.no_grad():
        ts = torch.as_tensor(np.asarray(data_ts), dtype=torch.float32)
        init_ts = torch.cat([ts[1:2], ts[0:1]]).to(dev)
        y0 = torch.as_tensor(np.asarray(y_first), dtype=torch.float32).to(dev)
        x0s = y0.unsqueeze(0).expand(num_samples, *y0.shape).contiguous()
        return model.forward_sets(x0s, init_ts, rng=rng)[:, :, -1].mean(0)


def initialize_latents_with_data(model, data_ys, data_ts, num_samples=None, rng="numpy"):
    """q(x0) mean <- backward-in-time solve from the first observation (20 draws, ``src/gpode/model_initialization.py:
    55-76``); for the shooting model also the shooting-state means <- the observations ``data_ys[:, :-1]`` (50
    draws, ``src/gpode_shooting/model_initialization.py:55-76``)."""
    shooting = not hasattr(model, "x0_distribution")
    if num_samples is None:
        num_samples = 50 if shooting else 20
    data_ys = np.asarray(data_ys)
    init_xs = data_ys[:, :-1]
    init_x0 = initial_state_from_data(model, init_xs[:, 0], data_ts, num_samples, rng=rng)
    with torch.no_grad():
        if shooting:
            model.state_distribution._initialize(init_x0, torch.as_tensor(init_xs, dtype=torch.float32))
        else:
            model.x0_distribution._initialize(init_x0)
    return model


def initialize_noisevar(model, init_noisevar):
    """Observation-noise variance of the likelihood (reference ``src/gpode_shooting/model_initialization.py:79-89``)."""
    p = model.likelihood.unconstrained_variance
    with torch.no_grad():
        p.copy_(constraint_utils.invsoftplus(torch.as_tensor(init_noisevar, dtype=p.dtype, device=p.device)
                                             * torch.ones_like(p)))
    return model


def initialize_and_fix_kernel_parameters(model, lengthscale_value=1.25, variance_value=0.5, fix=False):
    """Constant kernel hyper-parameters, optionally frozen (reference ``model_initialization.py:92-111``)."""
    kern = model.flow.odefunc.diffeq.kern
    with torch.no_grad():
        kern.unconstrained_lengthscales.copy_(constraint_utils.invsoftplus(
            lengthscale_value * torch.ones_like(kern.unconstrained_lengthscales)))
        kern.unconstrained_variance.copy_(constraint_utils.invsoftplus(
            variance_value * torch.ones_like(kern.unconstrained_variance)))
    if fix:
        kern.unconstrained_lengthscales.requires_grad_(False)
        kern.unconstrained_variance.requires_grad_(False)
    return model
