"""Variational posteriors of the initial state q(x0) and of the shooting states q(s_1..T) (mirror of reference
``src/core/states.py``; same class names, parameter names and shapes).

ELBO side terms either side of the integrator: on a CUDA device
N(m, L L^T + 1e-5 I) is sampled and its entropy taken by the fused ``gpode_state_fwd/_bwd`` kernels straight from the
packed lower-triangular parameters (one thread per D x D matrix); the spelled-out tensor algebra below -- one batched
Cholesky, what ``MultivariateNormal`` does internally without its host-synchronising argument validation, and the
closed-form entropy 0.5 D (1 + log 2 pi) + sum log diag -- serves host-side checks and D > 8."""
import math

import numpy as np
import torch
from torch import nn

from .. import ops
from ..misc import transforms
from ..misc.param import Param
from ..misc.torch_utils import host_to_device

initial_state_scale = 1e-1
jitter = 1e-5


def sample_normal(shape, seed=None):
    rng = np.random if seed is None else np.random.RandomState(seed)
    return torch.tensor(rng.normal(size=shape).astype(np.float32))


def _standard_normal(shape, dtype, device):
    """The reparameterisation noise of ``rsample`` (torch global generator, like ``MultivariateNormal.rsample``);
    a module-level function so tests can inject draws."""
    return torch.randn(shape, dtype=dtype, device=device)


class _FullRankGaussian:
    """N(mean, lchol lchol^T + jitter I) over the last axis, batched over the leading axes.

    On a CUDA device sampling and entropy run in the fused kernels of ``gpode_state_fwd/_bwd`` straight from the
    PACKED lower-triangular parameter (one thread per D x D matrix: Cholesky, matrix-vector product, log-determinant
    and their backward in registers); elsewhere the same algebra is spelled out with torch ops."""

    def __init__(self, mean, lchol_fn, packed):
        self.loc = mean
        self._lchol_fn = lchol_fn
        self._packed = packed
        self._fused = mean.is_cuda and mean.shape[-1] <= 8
        self._scale_tril = None

    @property
    def scale_tril(self):
        if self._scale_tril is None:
            lchol = self._lchol_fn()
            d = self.loc.shape[-1]
            cov = lchol @ lchol.transpose(-1, -2) + torch.eye(d, dtype=lchol.dtype, device=lchol.device) * jitter
            self._scale_tril = torch.linalg.cholesky(cov)
        return self._scale_tril

    def rsample(self, sample_shape, eps=None):
        """``eps``: the reparameterisation noise, if the caller has already drawn it (time sharding slices one draw)."""
        if eps is None:
            shape = tuple(sample_shape) + tuple(self.loc.shape)
            eps = _standard_normal(shape, self.loc.dtype, self.loc.device)
        if self._fused:
            return ops.state_sample(self.loc, self._packed, eps, jitter)
        return self.loc + (self.scale_tril @ eps.unsqueeze(-1)).squeeze(-1)

    def entropy(self):
        d = self.loc.shape[-1]
        if self._fused:
            return ops.state_entropy(self._packed, d, jitter)
        half_log_det = self.scale_tril.diagonal(dim1=-2, dim2=-1).log().sum(-1)
        return 0.5 * d * (1.0 + math.log(2 * math.pi)) + half_log_det

    def log_prob(self, x):
        d = self.loc.shape[-1]
        diff = (x - self.loc).unsqueeze(-1)
        tril = self.scale_tril
        z = torch.linalg.solve_triangular(tril.expand(diff.shape[:-2] + tril.shape[-2:]), diff, upper=False).squeeze(-1)
        half_log_det = tril.diagonal(dim1=-2, dim2=-1).log().sum(-1)
        return -0.5 * (d * math.log(2 * math.pi) + z.pow(2).sum(-1)) - half_log_det


class StateInitialDistribution(nn.Module):
    def __init__(self, dim_n, dim_d):
        super().__init__()
        self.dim_n = dim_n
        self.dim_d = dim_d

    def _initialize(self, x):
        raise NotImplementedError

    def sample(self, num_samples, seed=None):
        raise NotImplementedError

    def log_prob(self, x):
        raise NotImplementedError

    def kl(self):
        raise NotImplementedError


class StateInitialVariationalGaussian(StateInitialDistribution):
    """q(x0) = N(m, S), m (N,D), S (N,D,D) full rank (reference ``states.py:46-114``)."""

    def __init__(self, dim_n, dim_d):
        super().__init__(dim_n, dim_d)
        self.param_mean = Param(np.random.normal(size=(dim_n, dim_d)) * 1e-2, name='Initial state distribution (mean)')
        self.param_lchol = Param(np.stack([np.eye(dim_d)] * dim_n) * initial_state_scale,
                                 transform=transforms.LowerTriangular(dim_d, dim_n),
                                 name='Initial state distribution (scale)')

    def _initialize(self, x):
        self.param_mean.optvar.data = x.to(self.param_mean.optvar)

    def mean(self):
        return self.param_mean()

    def lchol(self):
        return self.param_lchol()

    def distribution(self):
        return _FullRankGaussian(self.mean(), self.lchol, self.param_lchol.optvar)

    def sample_numpy(self, num_samples=1, seed=None):
        eps = host_to_device(sample_normal(shape=(num_samples, self.dim_n, self.dim_d), seed=seed),
                             self.param_mean.optvar.device)
        return torch.einsum('nij, snj -> sni', self.lchol(), eps) + self.mean().unsqueeze(0)

    def sample(self, num_samples=1, seed=None):
        return self.distribution().rsample((num_samples,))  # (S,N,D)

    def log_prob(self, x):
        return self.distribution().log_prob(x)

    def kl(self):
        """KL[q(x0) || N(0, I)], summed over sequences (reference ``states.py:97-114``)."""
        alpha = self.mean()
        if alpha.is_cuda:
            # the same closed form as the whitened inducing KL -- 0.5 (sum alpha^2 + sum L^2 - sum log diag(L)^2 - count)
            # over N factors of size D x D -- so the fused kernel pair gpode_kl_fwd/_bwd serves it straight from the
            # packed parameter (two launches instead of ~27 element-wise ones; the sums are layout-independent, so the
            # (N, D) mean is passed as the kernel's (D, N) view without a copy)
            return ops.whitened_kl(alpha.reshape(self.dim_d, self.dim_n), self.param_lchol.optvar)
        Lq = torch.tril(self.lchol())
        Lq_diag = torch.diagonal(Lq, dim1=1, dim2=2)
        two_kl = (-torch.log(Lq_diag.pow(2)).sum(1) + alpha.pow(2).sum(1) + Lq.pow(2).sum(dim=(1, 2))
                  - float(self.dim_d))
        return 0.5 * two_kl.sum()


class StateSequenceVariationalDistribution(nn.Module):
    def __init__(self, dim_n, dim_t, dim_d):
        super().__init__()
        self.dim_n = dim_n
        self.dim_t = dim_t
        self.dim_d = dim_d

    def sample(self, num_samples, **kwargs):
        raise NotImplementedError

    def log_prob(self, x):
        raise NotImplementedError

    def entropy(self):
        raise NotImplementedError


class StateSequenceVariationalFactorizedGaussian(StateSequenceVariationalDistribution):
    """q(s) = prod_{n,t} N(m_nt, S_nt), plus q(x0) as ``self.x0`` (reference ``states.py:144-207``)."""

    def __init__(self, dim_n, dim_t, dim_d):
        super().__init__(dim_n, dim_t, dim_d)
        self.param_mean = Param(np.random.normal(size=(dim_n, dim_t, dim_d)) * 1e-1, name='State distribution (mean)')
        self.param_lchol = Param(np.stack([np.stack([np.eye(dim_d)] * dim_t)] * dim_n) * initial_state_scale,
                                 transform=transforms.StackedLowerTriangular(dim_d, dim_n, dim_t),
                                 name='State distribution (scale)')
        self._add_initial_state()

    def _add_initial_state(self):
        self.x0 = StateInitialVariationalGaussian(self.dim_n, self.dim_d)

    def _initialize(self, x0, xs, xs_std=None):
        self.x0._initialize(x0)
        self.param_mean.optvar.data = xs.to(self.param_mean.optvar)
        if xs_std is not None:
            self.param_lchol.optvar.data = self.param_lchol.transform.backward_tensor(
                torch.diag_embed(xs_std)).to(self.param_lchol.optvar)

    def mean(self):
        return self.param_mean()

    def lchol(self):
        return self.param_lchol()

    def distribution(self):
        return _FullRankGaussian(self.mean(), self.lchol, self.param_lchol.optvar)

    def sample_numpy(self, num_samples=1, seed=None):
        dev = self.param_mean.optvar.device
        eps = host_to_device(sample_normal(shape=(num_samples, self.dim_n, self.dim_t, self.dim_d), seed=seed), dev)
        zs = torch.einsum('ntij, sntj->snti', self.lchol(), eps)
        return torch.cat([self.x0.sample(num_samples, seed).unsqueeze(2), zs + self.mean().unsqueeze(0)], 2)

    def sample(self, num_samples=1, seed=None):
        """(S, N, T+1, D): the x0 sample followed by the T shooting states (reference ``states.py:199-201``)."""
        return torch.cat([self.x0.sample(num_samples).unsqueeze(2),
                          self.distribution().rsample((num_samples,))], 2)

    def entropy(self):
        return self.distribution().entropy()  # (N,T)

    # ---- time sharding (distributed.enable_time_sharding): only a slice of the time axis on this process ----------
    def _slice_distribution(self, a, b):
        """q over the shooting states a..b-1 (views of the parameters: gradients flow to those rows only)."""
        lchol_fn = self.lchol
        return _FullRankGaussian(self.mean()[:, a:b].contiguous(), lambda: lchol_fn()[:, a:b],
                                 self.param_lchol.optvar[:, a:b].contiguous())

    def sample_time_slice(self, num_samples, lo, hi):
        """Time indices ``lo..hi-1`` of :meth:`sample` (index 0 = the x0 sample, index t >= 1 = shooting state t - 1):
        the SAME numbers -- the full reparameterisation noise is drawn in the same order and then sliced -- but only this
        slice's Cholesky factors and matrix-vector products are computed. -> ``(S, N, hi - lo, D)``."""
        S, N, T, D = num_samples, self.dim_n, self.dim_t, self.dim_d
        ref = self.param_mean.optvar
        eps0 = _standard_normal((S, N, D), ref.dtype, ref.device)
        eps = _standard_normal((S, N, T, D), ref.dtype, ref.device)
        parts = []
        if lo == 0:
            parts.append(self.x0.distribution().rsample((S,), eps=eps0).unsqueeze(2))
        a, b = max(lo, 1) - 1, hi - 1
        if b > a:
            parts.append(self._slice_distribution(a, b).rsample((S,), eps=eps[:, :, a:b].contiguous()))
        return torch.cat(parts, 2) if len(parts) > 1 else parts[0]

    def entropy_time_slice(self, lo, hi):
        """Entropy of the shooting states with time indices ``lo..hi-1`` (index 0, the initial state, has none)."""
        a, b = max(lo, 1) - 1, hi - 1
        if b <= a:
            return torch.zeros((self.dim_n, 0), dtype=self.param_mean.optvar.dtype, device=self.param_mean.optvar.device)
        return self._slice_distribution(a, b).entropy()

    def log_prob(self, x):
        return self.distribution().log_prob(x)
