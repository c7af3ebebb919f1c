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
