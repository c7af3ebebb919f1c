"""Drop-in for the two ``torchdiffeq`` entry points the reference imports (``src/core/flow.py:3-4``):

    odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None) -> (len(t), *y0.shape)
    odeint_adjoint(...)                                                   same arguments

``func`` must be a GP vector field: a ``DSVGP_Layer`` whose cache is built, or an ``ODEfunc`` wrapping one. Then the
whole integration -- every stage of every step -- runs inside ONE persistent CUDA kernel (``gpode_rk4_fwd`` for
``method='rk4'``, ``gpode_dopri5_fwd`` for ``'dopri5'``), differentiable through the hand-written discrete adjoint.
Anything else raises: this library accelerates the GPODE path only and carries no generic / CPU fallback solver.
"""
import torch

from . import ops
from ._lib import GpodeError

FUSED_METHODS = ("rk4", "dopri5")


def _resolve(func):
    """-> (layer, counter) where ``layer`` is the DSVGP_Layer and ``counter`` an object with ``count_evals``."""
    from .core.dsvgp import DSVGP_Layer
    layer = getattr(func, "diffeq", func)
    if not isinstance(layer, DSVGP_Layer):
        raise GpodeError("odeint: func must be a DSVGP_Layer or an ODEfunc wrapping one, got %s "
                         "(no generic solver / fallback exists in this library)" % type(func).__name__)
    if not hasattr(layer, "nu"):
        raise GpodeError("odeint: the layer has no cache; call build_cache() (Flow.forward does) before integrating")
    return layer, (func if hasattr(func, "count_evals") else None)


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    if isinstance(y0, (tuple, list)):
        raise NotImplementedError("tuple states (the reference's dead divergence branch) are not supported")
    method = "dopri5" if method is None else method
    if method not in FUSED_METHODS:
        raise GpodeError("odeint: method %r has no fused CUDA integrator (available: %s)" % (method, FUSED_METHODS))
    layer, counter = _resolve(func)
    if t.ndim != 1 or t.numel() < 1:
        raise GpodeError("odeint: t must be a 1-D tensor with at least one point")
    t = t.to(y0.device)
    if method == "rk4":
        xs = ops.rk4_integrate(y0, t.to(torch.float32), *layer.cache_tensors())
        nfe = 4 * (t.numel() - 1)
    else:
        xs, stats = ops.dopri5_integrate(y0, t, *layer.cache_tensors(), rtol=rtol, atol=atol)
        nfe = stats[0]
    if counter is not None:
        counter.count_evals(nfe)
    return xs


def odeint_adjoint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None, **unused):
    """The reference never enables the continuous adjoint (``use_adjoint=False`` in all four scripts); requests for it
    are served by the same fused integrator and its discrete adjoint (exact gradients of the discrete solve)."""
    return odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
