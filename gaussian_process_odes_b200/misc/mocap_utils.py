"""Latent (PCA) space -> observation space decoders (mirror of reference ``src/misc/mocap_utils.py:12-34``).

``LinearProjection`` is the fixed affine map ``x -> (x * scale + shift) @ components``; ``Latent2DataProjector`` builds
it from a dataset object exactly like the reference does. Both expose ``as_affine()`` so that the likelihood can fuse
decoder and Gaussian log-density into one kernel instead of materialising the decoded ``(S,N,T,D_obs)`` tensor."""
import numpy as np
import torch

from .settings import settings


class LinearProjection:
    def __init__(self, components, scale=None, shift=None, device=None):
        dev = settings.device if device is None else device
        as_t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dev)
        self.pca_components = as_t(components)  # (D_latent, D_obs)
        self.pca_normalize_std = as_t(scale)
        self.pca_normalize_mean = as_t(shift)
        self._affine = None

    def inverse_pca_normalization(self, x):
        if self.pca_normalize_std is None:
            return x
        return (x * self.pca_normalize_std) + self.pca_normalize_mean

    def inverse_pca(self, x):
        return x @ self.pca_components

    def __call__(self, x):
        return self.inverse_pca(self.inverse_pca_normalization(x))

    def as_affine(self):
        """(W, b) with ``self(x) == x @ W + b``."""
        if self._affine is None:
            W = self.pca_components
            b = None
            if self.pca_normalize_std is not None:
                b = self.pca_normalize_mean @ W
                W = self.pca_normalize_std.unsqueeze(1) * W
            self._affine = (W.contiguous(), None if b is None else b.contiguous())
        return self._affine


class Latent2DataProjector(LinearProjection):
    """Built from a dataset with ``pca.components_``, optional ``pca_normalize.mean/std`` and ``data_mean/std``."""

    def __init__(self, dataset):
        norm = getattr(dataset, "pca_normalize", None)
        super().__init__(dataset.pca.components_, None if norm is None else norm.std,
                         None if norm is None else norm.mean)
        dev = settings.device
        self.data_std = torch.as_tensor(np.asarray(dataset.data_std, dtype=np.float32)).to(dev)
        self.data_mean = torch.as_tensor(np.asarray(dataset.data_mean, dtype=np.float32)).to(dev)
