"""RBF kernel hyper-parameters and Fourier-frequency sampling (mirror of reference ``src/core/kernels.py``).

Parameter names (``unconstrained_lengthscales``, ``unconstrained_variance``), shapes and initial values are the
reference's (``kernels.py:33-51``). ``K`` / ``square_dist*`` are kept for API completeness as plain tensor algebra;
the hot path never calls them -- K(x,Z) inside the vector field and K(Z,Z) inside the whitening are computed by the
CUDA kernels directly from ``lengthscales`` / ``variance``."""
import numpy as np
import torch
from torch import nn
from torch.nn import init

from ..misc.constraint_utils import invsoftplus, softplus
from ..misc.settings import settings
from ..misc.torch_utils import host_to_device


def sample_normal(shape, seed=None):
    """Standard-normal draw on the host (reference ``kernels.py:13-15``). The reference builds a fresh UNSEEDED
    ``RandomState()`` here when ``seed is None``; this mirror draws from numpy's global generator instead so that
    ``seed_everything`` makes omega reproducible (same distribution)."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    return torch.tensor(rng.normal(size=shape).astype(np.float32))


class RBF(nn.Module):
    def __init__(self, D_in, D_out=None, dimwise=False):
        super().__init__()
        self.D_in = D_in
        self.D_out = D_in if D_out is None else D_out
        self.dimwise = dimwise
        ls_shape = (self.D_out, self.D_in) if dimwise else (self.D_in,)
        var_shape = (self.D_out,) if dimwise else (1,)
        dev = settings.device
        self.unconstrained_lengthscales = nn.Parameter(torch.ones(size=ls_shape, device=dev), requires_grad=True)
        self.unconstrained_variance = nn.Parameter(torch.ones(size=var_shape, device=dev), requires_grad=True)
        self._initialize()

    def _initialize(self):
        init.constant_(self.unconstrained_lengthscales, invsoftplus(torch.tensor(1.3)).item())
        init.constant_(self.unconstrained_variance, invsoftplus(torch.tensor(0.5)).item())

    @property
    def lengthscales(self):
        return softplus(self.unconstrained_lengthscales)

    @property
    def variance(self):
        return softplus(self.unconstrained_variance)

    # the (D_out, D_in) / (D_out,) view the CUDA kernels take, whatever ``dimwise`` is
    def lengthscales_dimwise(self):
        ls = self.lengthscales
        return ls if self.dimwise else ls.unsqueeze(0).expand(self.D_out, self.D_in)

    def variance_dimwise(self):
        v = self.variance
        return v if self.dimwise else v.expand(self.D_out)

    def square_dist_dimwise(self, X, X2=None):
        Xs = X.unsqueeze(0) / self.lengthscales.unsqueeze(1)
        X2s = Xs if X2 is None else X2.unsqueeze(0) / self.lengthscales.unsqueeze(1)
        return (Xs.unsqueeze(2) - X2s.unsqueeze(1)).pow(2).sum(-1)  # (D_out, N, M), direct form

    def square_dist(self, X, X2=None):
        Xs = X / self.lengthscales
        X2s = Xs if X2 is None else X2 / self.lengthscales
        return (Xs.unsqueeze(1) - X2s.unsqueeze(0)).pow(2).sum(-1)  # (N, M)

    def K(self, X, X2=None):
        if self.dimwise:
            return self.variance[:, None, None] * torch.exp(-0.5 * self.square_dist_dimwise(X, X2))
        return self.variance * torch.exp(-0.5 * self.square_dist(X, X2))

    def sample_freq(self, S, seed=None, lengthscales=None):
        """omega = eps / lengthscale, ``(D_in, S, D_out)`` if dimwise else ``(D_in, S)`` (``kernels.py:101-112``).
        ``lengthscales``: the already constrained parameter, if the caller holds it (one softplus per cache build)."""
        shape = (self.D_in, S, self.D_out) if self.dimwise else (self.D_in, S)
        eps = host_to_device(sample_normal(shape, seed), self.unconstrained_lengthscales.device)
        ell = self.lengthscales if lengthscales is None else lengthscales
        ls = ell.T.unsqueeze(1) if self.dimwise else ell.unsqueeze(1)
        return eps / ls
