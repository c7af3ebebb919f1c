"""Observation likelihoods (mirror of reference ``src/core/likelihoods.py``)."""
import numpy as np
import torch
import torch.nn as nn
from torch.nn import init

from ..misc.constraint_utils import invsoftplus, softplus
from ..misc.settings import settings


class Gaussian(nn.Module):
    """Independent Gaussian noise with a learnt variance per observed dimension (reference ``likelihoods.py:10-28``)."""

    def __init__(self, ndim=1, init_val=0.25):
        super().__init__()
        self.unconstrained_variance = torch.nn.Parameter(torch.ones(ndim, device=settings.device), requires_grad=True)
        self._initialize(init_val)

    def _initialize(self, x):
        init.constant_(self.unconstrained_variance, invsoftplus(torch.tensor(x)).item())

    @property
    def variance(self):
        return softplus(self.unconstrained_variance)

    def log_prob(self, F, Y):
        var = self.variance
        return -0.5 * (np.log(2.0 * np.pi) + torch.log(var) + torch.pow(F - Y, 2) / var)

    # ---- fused mean log-likelihood (what both ELBOs reduce log_prob to: models.py:56-58 / shooting models.py:128,143) ----
    def _affine(self, F):
        """(W (D,D_obs), bias or None) of the latent -> observation map, or None if it is not a known affine map."""
        d = F.shape[-1]
        eye = getattr(self, "_eye", None)
        if eye is None or eye.shape[0] != d or eye.device != F.device:
            eye = torch.eye(d, dtype=F.dtype, device=F.device)
            self._eye = eye
        return eye, None

    def log_prob_mean(self, F, Y):
        """``log_prob(F, Y).mean()`` in one kernel (value + gradients) when the shapes allow it."""
        aff = self._affine(F) if F.is_cuda else None
        if aff is not None:
            W, b = aff
            lead = tuple(F.shape[:-1])
            ylead = tuple(Y.shape[:-1])
            if Y.shape[-1] == W.shape[1] and W.shape[0] <= 8 and W.shape[1] <= 128 and (
                    ylead == lead or (len(ylead) == len(lead) and ylead[0] == 1 and ylead[1:] == lead[1:])):
                from .. import ops
                var = self.variance
                if var.numel() != W.shape[1]:  # e.g. the class default Gaussian(ndim=1): broadcast like log_prob does
                    var = var.expand(W.shape[1])
                return ops.loglik_mean(F, Y, W, b, var)
        return self.log_prob(F, Y).mean()


class ProjectedGaussian(Gaussian):
    """Gaussian likelihood behind a fixed latent->data projection (reference ``likelihoods.py:31-45``)."""

    def __init__(self, projection, ndim=1, init_val=0.25):
        super().__init__(ndim, init_val)
        self.projection = projection

    def log_prob(self, F, Y):
        if F.ndim == 4:
            F = torch.stack([self.projection(_F) for _F in F])
        else:
            F = self.projection(F)
        return super().log_prob(F, Y)

    def _affine(self, F):
        as_affine = getattr(self.projection, "as_affine", None)
        return as_affine() if as_affine is not None else None
