"""Shared helpers of the parity tests.

Tolerances are the ones BASELINE.json's north_star states: relative 1e-5 on vector-field values, 1e-4 on integrated
trajectories, 1e-4 on ELBO terms and gradients ("relative" = max-abs error over max-abs value of the tensor).
Where the reference's OWN float32 path is further than that from the float64 evaluation of the same formula
(|nu| ~ 1e2..4e2 makes sum_m var nu_m K_m cancel heavily), the float64 oracle arbitrates:
the CUDA result must be at least as close to float64 as ARBITER_SLACK x the reference's float32 result is.
"""
import os

import numpy as np
import torch

import gpode_oracle as O

TOL_VF, TOL_TRAJ, TOL_GRAD = 1e-5, 1e-4, 1e-4
ARBITER_SLACK = 1.5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def assert_parity(name, cuda, ref32, f64, tol):
    """cuda vs the float32 reference/oracle within tol, else arbitrated by float64."""
    e_direct = relerr(cuda, ref32)
    if e_direct <= tol:
        return e_direct
    e_cuda, e_ref = relerr(cuda, f64), relerr(ref32, f64)
    assert e_cuda <= max(tol, ARBITER_SLACK * e_ref), (
        "%s: cuda-vs-ref32 %.3e > tol %.1e and cuda-vs-fp64 %.3e > %.1f x ref32-vs-fp64 %.3e"
        % (name, e_direct, tol, e_cuda, ARBITER_SLACK, e_ref))
    return e_cuda


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    t = lambda a: torch.tensor(a)
    out = dict(p={k[5:]: t(z[k]) for k in z.files if k.startswith("in_p_")},
               draws={k[8:]: t(z[k]) for k in z.files if k.startswith("in_draw_")},
               ys=t(z["in_ys"]), ts=t(z["in_ts"]),
               ref={k[4:]: t(z[k]) for k in z.files if k.startswith("ref_")},
               f64={k[4:]: t(z[k]) for k in z.files if k.startswith("f64_")},
               meta=[str(s) for s in z["meta"]])
    out["proj"] = t(z["in_proj_components"]) if "in_proj_components" in z.files else None
    return out


def to_dev(tree, dev="cuda"):
    if isinstance(tree, dict):
        return {k: to_dev(v, dev) for k, v in tree.items()}
    return tree.to(dev) if torch.is_tensor(tree) else tree


def oracle_cache(p, draws, dtype=torch.float32):
    """constrained GP parameters + cache from the oracle port, in ``dtype`` on CPU"""
    gp = O.gp_params(O.cast(p, dtype))
    d = O.cast(draws, dtype)
    c = O.build_cache(gp['Z'], gp['Um'], gp['Us_sqrt'], gp['ell'], gp['var'], d['w'], d['eps_omega'], d['phase_u'],
                      d['eps_u'])
    return gp, c
