"""Constrained <-> unconstrained variable transforms (mirror of reference ``src/misc/transforms.py``).

Same classes, method names and packed layouts (row-major ``np.tril_indices`` order). The two lower-triangular
``forward_tensor`` methods are ONE vectorised scatter here instead of the reference's Python loops over matrices
(``transforms.py:70-76`` loops D times, ``:105-112`` loops N*T times -- ~25 ms per call at N=6,T=99 and N*T tiny
kernel launches on a GPU); the result is identical."""
import numpy as np
import torch
import torch.nn.functional as F

from .settings import settings


class Identity:
    def __str__(self):
        return 'Identity transformation'

    def forward_tensor(self, x):
        return x

    def backward_tensor(self, y):
        return y

    def forward(self, x):
        return x

    def backward(self, y):
        return y


class SoftPlus:
    def __init__(self, lower=1e-12):
        self._lower = lower

    def __str__(self):
        return 'Softplus transformation'

    def forward(self, x):
        return np.logaddexp(0, x) + self._lower

    def forward_tensor(self, x):
        return F.softplus(x) + self._lower

    def backward_tensor(self, y):
        ys = torch.max(y - self._lower, torch.tensor(torch.finfo(y.dtype).eps).to(y))
        return ys + torch.log(-torch.expm1(-ys))

    def backward(self, y):
        ys = np.maximum(y - self._lower, np.finfo(settings.numpy_float).eps)
        return ys + np.log(-np.expm1(-ys))


class _TrilScatter:
    """Shared vectorised ``(..., n(n+1)/2) <-> (..., n, n)`` scatter/gather for the two classes below."""

    def _init_indices(self, n):
        self.N = n
        self._rows, self._cols = np.tril_indices(n, 0)
        self._flat_cache = {}

    def _flat_index(self, device):
        key = str(device)
        if key not in self._flat_cache:
            self._flat_cache[key] = torch.as_tensor(self._rows * self.N + self._cols, dtype=torch.long, device=device)
        return self._flat_cache[key]

    def _scatter(self, x, lead_shape):
        out = torch.zeros(lead_shape + (self.N * self.N,), dtype=x.dtype, device=x.device)
        out[..., self._flat_index(x.device)] = x
        return out.reshape(lead_shape + (self.N, self.N))

    def _scatter_np(self, x, lead_shape):
        out = np.zeros(lead_shape + (self.N, self.N), dtype=settings.numpy_float)
        out[..., self._rows, self._cols] = x
        return out


class LowerTriangular(_TrilScatter):
    def __init__(self, N, num_matrices=1):
        self._init_indices(N)
        self.num_matrices = num_matrices

    def __str__(self):
        return 'Lower cholesky transformation'

    def forward(self, x):
        return self._scatter_np(np.asarray(x), (self.num_matrices,))

    def backward(self, y):
        return np.asarray(y)[..., self._rows, self._cols].reshape(len(y), -1)

    def forward_tensor(self, x):
        return self._scatter(x, (self.num_matrices,))

    def backward_tensor(self, y):
        return y.reshape(y.shape[:-2] + (self.N * self.N,))[..., self._flat_index(y.device)]


class StackedLowerTriangular(_TrilScatter):
    def __init__(self, N, num_n, num_m):
        self._init_indices(N)
        self.num_n = num_n
        self.num_m = num_m

    def __str__(self):
        return 'Lower cholesky transformation for stack sequence of covariance matrices'

    def forward(self, x):
        return self._scatter_np(np.asarray(x), (self.num_n, self.num_m))

    def backward(self, y):
        return np.asarray(y)[..., self._rows, self._cols]

    def forward_tensor(self, x):
        return self._scatter(x, (self.num_n, self.num_m))

    def backward_tensor(self, y):
        return y.reshape(y.shape[:-2] + (self.N * self.N,))[..., self._flat_index(y.device)]
