"""Model assembly and loss functions (mirror of reference ``src/gpode/model_builder.py:18-57`` and
``src/gpode_shooting/model_builder.py:19-72`` / ``mocap_model_builder.py``): ``build_model``-style helpers taking plain
keyword arguments instead of an argparse namespace, and the two ``compute_loss`` functions."""
import torch

from .core import constraints as constraints
from .core.dsvgp import DSVGP_Layer
from .core.flow import Flow
from .core.likelihoods import Gaussian, ProjectedGaussian
from .core.states import StateInitialVariationalGaussian, StateSequenceVariationalFactorizedGaussian
from .gpode.models import SequenceModel
from .gpode_shooting.models import UniformSequenceModel


def _likelihood(D_obs, projection):
    return Gaussian(ndim=D_obs) if projection is None else ProjectedGaussian(projection=projection, ndim=D_obs)


def build_gpode(N, T, D, num_inducing=16, num_features=256, solver='dopri5', ts_dense_scale=4, use_adjoint=False,
                dimwise=True, q_diag=False, D_obs=None, projection=None):
    gp = DSVGP_Layer(D_in=D, D_out=D, M=num_inducing, S=num_features, dimwise=dimwise, q_diag=q_diag)
    flow = Flow(diffeq=gp, solver=solver, use_adjoint=use_adjoint)
    D_obs = D if D_obs is None else D_obs
    return SequenceModel(flow=flow, num_observations=N * T * D_obs,
                         x0_distribution=StateInitialVariationalGaussian(dim_n=N, dim_d=D),
                         likelihood=_likelihood(D_obs, projection), ts_dense_scale=ts_dense_scale)


def build_gpode_shooting(N, T, D, num_inducing=16, num_features=256, solver='dopri5', ts_dense_scale=4,
                         use_adjoint=False, dimwise=True, q_diag=False, constraint_type='gauss',
                         constraint_initial_scale=1e-3, constraint_trainable=False, D_obs=None, projection=None):
    gp = DSVGP_Layer(D_in=D, D_out=D, M=num_inducing, S=num_features, dimwise=dimwise, q_diag=q_diag)
    flow = Flow(diffeq=gp, solver=solver, use_adjoint=use_adjoint)
    if constraint_type == 'gauss':
        constraint = constraints.Gaussian(d=1, scale=constraint_initial_scale, requires_grad=constraint_trainable)
    elif constraint_type == 'laplace':
        constraint = constraints.Laplace(d=1, scale=constraint_initial_scale, requires_grad=constraint_trainable)
    else:
        raise ValueError("invalid constraint likelihood specification, only available options are gauss/laplace")
    D_obs = D if D_obs is None else D_obs
    return UniformSequenceModel(
        flow=flow, num_observations=N * T * D_obs,
        state_distribution=StateSequenceVariationalFactorizedGaussian(dim_n=N, dim_t=T - 1, dim_d=D),
        likelihood=_likelihood(D_obs, projection), constraint=constraint, ts_dense_scale=ts_dense_scale)


def compute_loss_gpode(model, ys, ts):
    """-> loss, nll, initial_state_kl, inducing_kl (reference ``src/gpode/model_builder.py:46-57``)."""
    observ_loglik, init_state_kl = model.build_lowerbound_terms(ys, ts)
    kl = model.build_kl()
    loss = -(observ_loglik - init_state_kl - kl)
    return loss, -observ_loglik, init_state_kl, kl


def compute_loss_shooting(model, ys, ts, **kwargs):
    """-> loss, nll, state term, initial_state_kl, inducing_kl (``src/gpode_shooting/model_builder.py:59-72``)."""
    ll, cons, ent, k0 = model.build_lowerbound_terms(ys, ts, **kwargs)
    inducing_kl = model.build_inducing_kl()
    loss = -(ll + cons + ent - k0 - inducing_kl)
    return loss, -ll, -(cons + ent), k0, inducing_kl


def _x0_posterior(model):
    return model.x0_distribution if hasattr(model, "x0_distribution") else model.state_distribution.x0


def compute_predictions(model, ts, eval_sample_size=10, x0_distribution=None, batched=True, rng="numpy"):
    """Posterior-predictive trajectories ``(S,N,T,D)`` from the optimised initial-state posterior, one GP function
    draw per sample (reference ``src/gpode/model_builder.py:60-78``, ``src/gpode_shooting/model_builder.py:75-93``).

    ``batched=True`` (default): all ``eval_sample_size`` draws go through ONE whitening, one pack and one integrator
    launch (``model.forward_sets``); the host generator is consumed in the reference's per-sample order
    (``rng='numpy'``) or skipped (``rng='device'``). ``batched=False`` is the reference's loop, one launch set per
    sample."""
    from .misc.torch_utils import insert_zero_t0
    model.eval()
    dist = _x0_posterior(model) if x0_distribution is None else x0_distribution
    ts = insert_zero_t0(ts)
    with torch.no_grad():
        if batched:
            x0 = dist.sample(num_samples=eval_sample_size)  # (S,N,D), one draw of the initial state per sample
            return model.forward_sets(x0, ts, rng=rng)[:, :, 1:]
        out = [model(dist.sample().squeeze(0), ts) for _ in range(eval_sample_size)]
    return torch.stack(out, 0)[:, :, 1:]


def compute_test_predictions(model, x0, ts, eval_sample_size=10, batched=True, rng="numpy"):
    """Predictive trajectories ``(S,N,T,D)`` from a GIVEN initial state ``x0 (N,D)`` (reference
    ``src/gpode/model_builder.py:81-96``, ``src/gpode_shooting/mocap_model_builder.py:104-119``)."""
    model.eval()
    with torch.no_grad():
        if batched:
            x0s = x0.unsqueeze(0).expand(eval_sample_size, *x0.shape).contiguous()
            return model.forward_sets(x0s, ts, rng=rng)
        return torch.stack([model(x0, ts) for _ in range(eval_sample_size)], 0)
