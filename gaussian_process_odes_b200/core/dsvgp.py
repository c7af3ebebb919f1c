"""``DSVGP_Layer``: the sparse-variational GP vector field sampled pathwise (decoupled sampling), mirror of reference
``src/core/dsvgp.py``. Same constructor, attributes (``rff_weights``, ``rff_omega``, ``rff_phase``, ``nu``), methods
(``build_cache``, ``rff_forward``, ``sample_inducing``, ``forward``, ``kl``) and ``state_dict`` keys; the arithmetic
runs in the sm_100a kernels of ``libgpode_b200.so``:

* ``build_cache``  -> ``gpode_whiten_fwd/bwd`` (Kzz, in-shared-memory Cholesky, two triangular solves, RFF at Z),
* ``forward(t,x)`` -> ``gpode_vf_fwd/bwd``,
* ``kl``           -> ``gpode_kl_fwd/bwd``.

There is no CPU fallback: on a machine without the CUDA library / a GPU these methods raise ``GpodeError``."""
import numpy as np
import torch

from .. import ops
from ..misc import transforms
from ..misc.param import Param
from ..misc.torch_utils import host_to_device
from .kernels import RBF

jitter = 1e-5


def sample_normal(shape, seed=None):
    """Host standard-normal draw, numpy global generator unless seeded (reference ``dsvgp.py:11-17``)."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    return torch.tensor(rng.normal(size=shape).astype(np.float32))


def sample_uniform(shape, seed=None):
    """Host U(0,1) draw (reference ``dsvgp.py:20-26``)."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    return torch.tensor(rng.uniform(low=0.0, high=1.0, size=shape).astype(np.float32))


class DSVGP_Layer(torch.nn.Module):
    def __init__(self, D_in, D_out, M, S, q_diag=False, dimwise=True):
        super().__init__()
        if D_in != D_out:
            raise ValueError("the ODE vector field needs D_in == D_out (got %d, %d)" % (D_in, D_out))
        self.kern = RBF(D_in, D_out, dimwise)
        self.q_diag = q_diag
        self.dimwise = dimwise
        self.D_out = D_out
        self.D_in = D_in
        self.M = M
        self.S = S
        self.inducing_loc = Param(np.random.normal(size=(M, D_in)), name='Inducing locations')
        self.Um = Param(np.random.normal(size=(M, D_out)) * 1e-1, name='Inducing distribution (mean)')
        if self.q_diag:
            self.Us_sqrt = Param(np.ones(shape=(M, D_out)) * 1e-3, transform=transforms.SoftPlus(),
                                 name='Inducing distribution (scale)')
        else:
            self.Us_sqrt = Param(np.stack([np.eye(M)] * D_out) * 1e-3,
                                 transform=transforms.LowerTriangular(M, D_out),
                                 name='Inducing distribution (scale)')

    @property
    def _device(self):
        return self.inducing_loc.optvar.device

    def sample_inducing(self):
        """u ~ q(u) = N(Um, Us Us^T) in whitened coordinates, ``(M, D_out)`` (reference ``dsvgp.py:78-90``)."""
        epsilon = host_to_device(sample_normal(shape=(self.M, self.D_out), seed=None), self._device)
        if self.q_diag:
            ZS = self.Us_sqrt() * epsilon
        elif epsilon.is_cuda:
            # one kernel straight from the packed factor instead of tril scatter + batched product (and three more
            # launches in the backward)
            return ops.inducing_sample(self.Um(), self.Us_sqrt.optvar, epsilon)
        else:
            ZS = torch.einsum('dnm, md->nd', self.Us_sqrt(), epsilon)
        return ZS + self.Um()

    # dimwise views of the cache for the kernels (a non-dimwise layer is the dimwise one with shared columns)
    def _omega_dimwise(self):
        return self.rff_omega if self.dimwise else self.rff_omega.unsqueeze(2).expand(self.D_in, self.S, self.D_out)

    def _phase_dimwise(self):
        return self.rff_phase if self.dimwise else self.rff_phase.unsqueeze(2).expand(1, self.S, self.D_out)

    def build_cache(self):
        """Fix one function draw: Fourier features, an inducing sample and nu = Kzz^-1 (u - f_prior(Z)) in whitened
        form (reference ``dsvgp.py:92-122``). Draw order on the host RNG is the reference's: weights, omega, phase,
        epsilon."""
        dev = self._device
        # the constrained hyper-parameters are evaluated ONCE per cache build and belong to the
        # cache (``cache_tensors`` hands the same tensors to the integrator): one softplus + one backward per parameter
        # and step instead of one per access (ncu launch list of a VDP shooting step: 12 softplus launches)
        ell = self.kern.lengthscales
        self._ell_dimwise = ell if self.dimwise else ell.unsqueeze(0).expand(self.D_out, self.D_in)
        self._var_dimwise = self.kern.variance_dimwise()
        self.rff_weights = host_to_device(sample_normal((self.S, self.D_out)), dev)
        self.rff_omega = self.kern.sample_freq(self.S, lengthscales=ell)
        phase_shape = (1, self.S, self.D_out) if self.dimwise else (1, self.S)
        self.rff_phase = host_to_device(sample_uniform(phase_shape), dev) * 2 * np.pi
        inducing_val = self.sample_inducing()
        nu = ops.whiten(self.inducing_loc(), self._ell_dimwise, self._var_dimwise,
                        inducing_val, self._omega_dimwise(), self._phase_dimwise(), self.rff_weights, jitter)
        # reference shapes: (D,M,1) dimwise, (M,D) otherwise
        self.nu = nu.unsqueeze(2) if self.dimwise else nu.t()

    def build_cache_sets(self, n, rng="numpy"):
        """``n`` independent function draws at once for batched Monte-Carlo prediction (no gradient): the tensors of
        ``build_cache`` with a leading draw axis, in the dimwise layout of the kernels --
        ``(nu (n,D,M), omega (n,D,S,D), phase (n,S,D), w (n,S,D))``.

        ``rng='numpy'`` consumes the host generators draw by draw in the reference's order (weights, omega, phase,
        epsilon -- ``dsvgp.py:100-103,83``), i.e. exactly the numbers ``n`` successive ``build_cache()`` calls would
        use; ``rng='device'`` draws with torch's CUDA generator instead (same distributions, no host work)."""
        dev, D, S, M = self._device, self.D_out, self.S, self.M
        om_shape = (D, S, D) if self.dimwise else (D, S)
        ph_shape = (1, S, D) if self.dimwise else (1, S)
        if rng == "numpy":
            from . import kernels as _kernels
            ws, oms, phs, eps = [], [], [], []
            for _ in range(n):  # host order per draw; module-level samplers so tests can inject
                ws.append(sample_normal((S, D)))
                oms.append(_kernels.sample_normal(om_shape))
                phs.append(sample_uniform(ph_shape))
                eps.append(sample_normal((M, D)))
            w, om_eps, ph, ep = (torch.stack(v).to(dev, non_blocking=True) for v in (ws, oms, phs, eps))
        elif rng == "device":
            w = torch.randn(n, S, D, device=dev)
            om_eps = torch.randn(n, *om_shape, device=dev)
            ph = torch.rand(n, *ph_shape, device=dev)
            ep = torch.randn(n, M, D, device=dev)
        else:
            raise ValueError("rng must be 'numpy' or 'device'")
        with torch.no_grad():
            ls = self.kern.lengthscales
            omega = om_eps / (ls.T.unsqueeze(1) if self.dimwise else ls.unsqueeze(1))
            phase = (ph * 2 * np.pi).reshape(n, S, -1)
            if not self.dimwise:
                omega = omega.unsqueeze(3).expand(n, D, S, D)
                phase = phase.expand(n, S, D)
            omega, phase = omega.contiguous(), phase.contiguous()
            if self.q_diag:
                u = self.Us_sqrt() * ep + self.Um()
            else:
                u = torch.einsum('dnm,smd->snd', self.Us_sqrt(), ep) + self.Um()
            nu = ops.whiten_sets(self.inducing_loc(), self.kern.lengthscales_dimwise(), self.kern.variance_dimwise(),
                                 u.contiguous(), omega, phase, w, jitter)
        return nu, omega, phase, w

    def _nu_dimwise(self):
        return self.nu.squeeze(2) if self.dimwise else self.nu.t()

    def rff_forward(self, x):
        """Prior sample through the random Fourier features only (reference ``dsvgp.py:124-137``): the vector field
        with the pathwise update switched off."""
        zero_nu = torch.zeros(self.D_out, self.M, dtype=x.dtype, device=x.device)
        return ops.vector_field(x, self.inducing_loc(), self.kern.lengthscales_dimwise(),
                                self.kern.variance_dimwise(), zero_nu, self._omega_dimwise(), self._phase_dimwise(),
                                self.rff_weights)

    def cache_tensors(self):
        """(Z, ell, var, nu, omega, phase, w) in the dimwise layout the integrator kernels take: the tensors of the
        last ``build_cache`` (nu and omega were computed from exactly these Z / lengthscales / variances)."""
        if getattr(self, "_ell_dimwise", None) is None:
            return (self.inducing_loc(), self.kern.lengthscales_dimwise(), self.kern.variance_dimwise(),
                    self._nu_dimwise(), self._omega_dimwise(), self._phase_dimwise(), self.rff_weights)
        return (self.inducing_loc(), self._ell_dimwise, self._var_dimwise, self._nu_dimwise(), self._omega_dimwise(),
                self._phase_dimwise(), self.rff_weights)

    def forward(self, t, x):
        """f(x) for the cached function draw; ``t`` is ignored (autonomous ODE) (reference ``dsvgp.py:172-197``)."""
        return ops.vector_field(x, *self.cache_tensors())

    def kl(self):
        """KL[q(u) || N(0, I)] in whitened form (reference ``dsvgp.py:199-230``)."""
        if self.q_diag:
            alpha, Lq = self.Um(), self.Us_sqrt()
            two_kl = -torch.log(Lq.pow(2)).sum(0) + alpha.pow(2).sum(0) + Lq.pow(2).sum(0) - float(self.M)
            return 0.5 * two_kl.sum()
        return ops.whitened_kl(self.Um(), self.Us_sqrt.optvar)
