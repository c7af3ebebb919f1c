"""Shooting-constraint priors p(s_{t+1} | x(t+1; s_t)) (mirror of reference ``src/core/constraints.py``)."""
import math

import torch
import torch.nn as nn
from torch.nn import init

from ..misc.constraint_utils import invsoftplus, softplus
from ..misc.settings import settings


class _ScaleFamily(nn.Module):
    def __init__(self, d=1, scale=1.0, requires_grad=True):
        super().__init__()
        self.unconstrained_scale = torch.nn.Parameter(torch.ones(d, device=settings.device),
                                                      requires_grad=requires_grad)
        self._initialize(scale)

    def _initialize(self, x):
        init.constant_(self.unconstrained_scale, invsoftplus(torch.tensor(x)).item())

    @property
    def scale(self):
        return softplus(self.unconstrained_scale)


class Gaussian(_ScaleFamily):
    """N(y; f, scale^2) elementwise (reference ``constraints.py:9-36``)."""

    @property
    def variance(self):
        return self.scale.pow(2)

    def log_prob(self, f, y):
        s = self.scale
        out = -(y - f).pow(2) / (2 * s.pow(2)) - s.log() - 0.5 * math.log(2 * math.pi)
        assert out.shape == f.shape
        return out


class Laplace(_ScaleFamily):
    """Laplace(y; f, scale) elementwise (reference ``constraints.py:39-66``)."""

    @property
    def variance(self):
        return 2 * self.scale.pow(2)

    def log_prob(self, f, y):
        s = self.scale
        out = -torch.log(2 * s) - torch.abs(y - f) / s
        assert out.shape == f.shape
        return out
