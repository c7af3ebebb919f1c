"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference modules on CPU: from ``/root/reference`` in the build
container, from the archive ``oracle/_ref/gpode_reference_src.zip`` (``oracle/stage_reference.py``) on the GPU box. It
* puts ``oracle/torchdiffeq_shim`` on ``sys.path`` so reference ``src/core/flow.py:3-4`` imports unchanged,
* replaces the reference's three numpy RNG helpers (``src/core/dsvgp.py:11-26``, ``src/core/kernels.py:13-15`` --
  the latter builds an UNSEEDED ``RandomState()``, so omega is irreproducible without this) and torch's
  ``_standard_normal`` used by ``MultivariateNormal.rsample`` (``src/core/states.py:91-92,199-201``) with queues
  of injected draws,
* builds the reference models from the same unconstrained-parameter dict the oracle port and the CUDA path use.
"""
import os
import sys
import warnings

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference tree itself (build container), else the byte-for-byte archive of its src/ staged by
# oracle/stage_reference.py (git-ignored; travels to the GPU box): either way the UNMODIFIED modules are imported
STAGED_ARCHIVE = os.path.join(_HERE, "_ref", "gpode_reference_src.zip")
REFERENCE_ROOT = os.environ.get("GPODE_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "core")) and os.path.isfile(STAGED_ARCHIVE):
    REFERENCE_ROOT = STAGED_ARCHIVE


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "core")) or os.path.isfile(REFERENCE_ROOT)


def source():
    """'tree' (the reference checkout), 'archive' (oracle/_ref) or None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "core")):
        return "tree"
    return "archive" if os.path.isfile(REFERENCE_ROOT) else None


def _import_reference():
    if not available():
        raise RuntimeError("reference not found: neither %s nor %s" % (REFERENCE_ROOT, STAGED_ARCHIVE))
    shim = os.path.join(_HERE, "torchdiffeq_shim")
    for p in (shim, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore")
    import src.core.dsvgp as dsvgp
    import src.core.kernels as kernels
    import src.core.flow as flow
    import src.core.states as states
    import src.core.likelihoods as likelihoods
    import src.core.constraints as constraints
    import src.gpode.models as gpode_models
    import src.gpode_shooting.models as shooting_models
    import torch.distributions.multivariate_normal as mvn
    # The reference's device singleton answers cuda:0 whenever a GPU is visible (src/misc/settings.py:18-19) while its
    # RNG helpers create CPU tensors: on the GPU box that mixes devices. The CPU arm pins the property
    # to the CPU at run time (nothing in the modules changes), as reference_in_float64 does for the dtype.
    import src.misc.settings as _rs
    type(_rs.settings).device = property(lambda self: torch.device('cpu'))
    return dict(dsvgp=dsvgp, kernels=kernels, flow=flow, states=states, likelihoods=likelihoods,
                constraints=constraints, gpode_models=gpode_models, shooting_models=shooting_models, mvn=mvn)


class _Queue:
    def __init__(self, items, what):
        self.items, self.what = list(items), what

    def __call__(self, shape, *a, **k):
        assert self.items, "reference asked for more %s draws than were injected" % self.what
        x = self.items.pop(0)
        assert tuple(x.shape) == tuple(shape), (self.what, tuple(x.shape), tuple(shape))
        return x.clone()


class injected_draws:
    """Context manager: the next ``build_cache`` / ``rsample`` calls of the reference consume ``draws``."""

    def __init__(self, mods, draws, n_caches=1, mvn_order=()):
        self.m, self.d, self.n, self.mvn_order = mods, draws, n_caches, mvn_order

    def __enter__(self):
        m, d = self.m, self.d
        self.saved = (m['dsvgp'].sample_normal, m['dsvgp'].sample_uniform, m['kernels'].sample_normal,
                      m['mvn']._standard_normal)
        # order inside DSVGP_Layer.build_cache: weights (S,D) [dsvgp.py:100], omega [kernels.py:109],
        # phase [dsvgp.py:103], epsilon (M,D) [dsvgp.py:83]
        m['dsvgp'].sample_normal = _Queue([d['w'], d['eps_u']] * self.n, "dsvgp.normal")
        m['dsvgp'].sample_uniform = _Queue([d['phase_u']] * self.n, "dsvgp.uniform")
        m['kernels'].sample_normal = _Queue([d['eps_omega']] * self.n, "kernels.normal")
        q = _Queue([d[k] for k in self.mvn_order], "mvn.standard_normal")
        m['mvn']._standard_normal = lambda shape, dtype, device: q(shape)
        return self

    def __exit__(self, *exc):
        m = self.m
        (m['dsvgp'].sample_normal, m['dsvgp'].sample_uniform, m['kernels'].sample_normal,
         m['mvn']._standard_normal) = self.saved
        return False


class reference_in_float64:
    """Evaluate the UNMODIFIED reference in float64: its global dtype singleton (src/misc/settings.py:21-27) and
    torch's default dtype are switched for the duration of the block. Used as the arbiter of the branches the oracle
    port does not restate (dimwise=False)."""

    def __enter__(self):
        _import_reference()
        import src.misc.settings as rs
        self.cls = type(rs.settings)
        self.saved = (self.cls.torch_float, self.cls.numpy_float, torch.get_default_dtype())
        self.cls.torch_float = property(lambda self: torch.float64)
        self.cls.numpy_float = property(lambda self: np.float64)
        torch.set_default_dtype(torch.float64)

    def __exit__(self, *exc):
        self.cls.torch_float, self.cls.numpy_float = self.saved[0], self.saved[1]
        torch.set_default_dtype(self.saved[2])
        return False


def _set(param, value):
    with torch.no_grad():
        param.copy_(value)


def build_reference_layer(mods, p, S):
    """A reference ``DSVGP_Layer`` (src/core/dsvgp.py:46-76) holding the unconstrained parameters ``p``."""
    M, D = p['inducing_loc'].shape
    gp = mods['dsvgp'].DSVGP_Layer(D_in=D, D_out=D, M=M, S=S, dimwise=True, q_diag=False)
    _set(gp.inducing_loc.optvar, p['inducing_loc'])
    _set(gp.Um.optvar, p['Um'])
    _set(gp.Us_sqrt.optvar, p['Us_sqrt_packed'])
    _set(gp.kern.unconstrained_lengthscales, p['unconstrained_lengthscales'])
    _set(gp.kern.unconstrained_variance, p['unconstrained_variance'])
    return gp


def build_reference_gpode(mods, p, ys, S, solver='rk4', ts_dense_scale=4, project=None, num_observations=None):
    """src/gpode/model_builder.py:18-43 with parameters overwritten by ``p``."""
    N, T, Dobs = ys.shape
    gp = build_reference_layer(mods, p, S)
    D = gp.D_in
    flow = mods['flow'].Flow(diffeq=gp, solver=solver, use_adjoint=False)
    if project is None:
        lik = mods['likelihoods'].Gaussian(ndim=Dobs)
    else:
        lik = mods['likelihoods'].ProjectedGaussian(projection=project, ndim=Dobs)
    x0d = mods['states'].StateInitialVariationalGaussian(dim_n=N, dim_d=D)
    model = mods['gpode_models'].SequenceModel(
        flow=flow, num_observations=(N * T * Dobs if num_observations is None else num_observations),
        x0_distribution=x0d, likelihood=lik, ts_dense_scale=ts_dense_scale)
    _set(x0d.param_mean.optvar, p['x0_mean'])
    _set(x0d.param_lchol.optvar, p['x0_lchol_packed'])
    _set(lik.unconstrained_variance, p['lik_unconstrained_variance'])
    return model


def build_reference_shooting(mods, p, ys, S, solver='rk4', project=None, num_observations=None):
    """src/gpode_shooting/model_builder.py:19-56 with parameters overwritten by ``p`` (Gaussian constraint)."""
    N, T, Dobs = ys.shape
    gp = build_reference_layer(mods, p, S)
    D = gp.D_in
    flow = mods['flow'].Flow(diffeq=gp, solver=solver, use_adjoint=False)
    if project is None:
        lik = mods['likelihoods'].Gaussian(ndim=Dobs)
    else:
        lik = mods['likelihoods'].ProjectedGaussian(projection=project, ndim=Dobs)
    cons = mods['constraints'].Gaussian(d=1, scale=1e-3, requires_grad=False)
    sd = mods['states'].StateSequenceVariationalFactorizedGaussian(dim_n=N, dim_t=T - 1, dim_d=D)
    model = mods['shooting_models'].UniformSequenceModel(
        flow=flow, num_observations=(N * T * Dobs if num_observations is None else num_observations),
        state_distribution=sd, likelihood=lik, constraint=cons, ts_dense_scale=4)
    _set(sd.x0.param_mean.optvar, p['x0_mean'])
    _set(sd.x0.param_lchol.optvar, p['x0_lchol_packed'])
    _set(sd.param_mean.optvar, p['state_mean'])
    _set(sd.param_lchol.optvar, p['state_lchol_packed'])
    _set(lik.unconstrained_variance, p['lik_unconstrained_variance'])
    _set(cons.unconstrained_scale, p['constraint_unconstrained_scale'])
    return model


def reference_gpode_loss(model, ys, ts):
    """src/gpode/model_builder.py:46-57"""
    ll, k0 = model.build_lowerbound_terms(ys, ts)
    kl = model.build_kl()
    return -(ll - k0 - kl), dict(observ_loglik=ll, init_state_kl=k0, inducing_kl=kl)


def reference_shooting_loss(model, ys, ts, num_samples):
    """src/gpode_shooting/model_builder.py:59-72"""
    ll, c, e, k0 = model.build_lowerbound_terms(ys, ts, num_samples=num_samples)
    kl = model.build_inducing_kl()
    return -(ll + c + e - k0 - kl), dict(observ_loglik=ll, constraint_loglik=c, state_entropy=e,
                                         init_state_kl=k0.reshape(()), inducing_kl=kl)


def reference_grads(model, kind):
    """Gradients keyed by the shared unconstrained-parameter names."""
    gp = model.flow.odefunc.diffeq
    g = dict(inducing_loc=gp.inducing_loc.optvar.grad, Um=gp.Um.optvar.grad, Us_sqrt_packed=gp.Us_sqrt.optvar.grad,
             unconstrained_lengthscales=gp.kern.unconstrained_lengthscales.grad,
             unconstrained_variance=gp.kern.unconstrained_variance.grad,
             lik_unconstrained_variance=model.likelihood.unconstrained_variance.grad)
    if kind == 'gpode':
        g.update(x0_mean=model.x0_distribution.param_mean.optvar.grad,
                 x0_lchol_packed=model.x0_distribution.param_lchol.optvar.grad)
    else:
        sd = model.state_distribution
        g.update(x0_mean=sd.x0.param_mean.optvar.grad, x0_lchol_packed=sd.x0.param_lchol.optvar.grad,
                 state_mean=sd.param_mean.optvar.grad, state_lchol_packed=sd.param_lchol.optvar.grad)
    return {k: v.detach().clone() for k, v in g.items()}
