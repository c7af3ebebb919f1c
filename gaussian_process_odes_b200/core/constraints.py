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


    _laplace = False

    def shooting_sum(self, ss, pred):
        """sum over everything of ``log_prob(ss[..., 1:, :], pred[..., :-1, :])`` -- the only way the multiple-shooting
        ELBO uses this prior (reference ``src/gpode_shooting/models.py:134-135,143``). One fused kernel when the scale
        is frozen (the reference default) and the tensors live on a GPU; plain tensor ops otherwise."""
        if ss.is_cuda and not self.unconstrained_scale.requires_grad and self.unconstrained_scale.numel() == 1:
            from .. import ops
            return ops.constraint_sum(ss, pred, self.scale.detach(), laplace=self._laplace)
        return self.log_prob(ss[..., 1:, :], pred[..., :-1, :]).sum()


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
    _laplace = True

    @property
    def variance(self):
        return 2 * self.scale.pow(2)

    def log_prob(self, f, y):
        s = self.scale
        out = -torch.log(2 * s) - torch.abs(y - f) / s
        assert out.shape == f.shape
        return out
