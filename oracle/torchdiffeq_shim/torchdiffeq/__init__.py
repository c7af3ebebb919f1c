"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the parts of ``torchdiffeq==0.2.0`` that GPODE calls.

The reference pins ``torchdiffeq==0.2.0`` (reference ``pyproject.toml:20``) and imports
``odeint`` / ``odeint_adjoint`` from it (reference ``src/core/flow.py:3-4``, call sites ``:84-90``).
The package is a third-party dependency that is NOT vendored under ``/root/reference`` and cannot be
installed offline, so its published algorithm is restated here from the 0.2.0 release:

* ``rk4``    = fixed-grid solver on the user's ``t`` grid, step function ``rk4_alt_step_func`` (3/8 rule),
* ``dopri5`` = Dormand-Prince 5(4) with the Shampine tableau, FSAL, float64 time / float32 state,
               whole-tensor RMS error norm, 4th-order dense-output interpolation,
* ``euler`` / ``midpoint`` fixed-grid solvers (cheap to restate, used by nothing on the hot path).

PARITY UNPINNED for this file: the reference holds no golden vectors at the ``odeint`` boundary
(SURVEY.md section 4 / 8c), and the real package is unavailable to check against.

Only ``tests/``, ``__graft_entry__.smoke()``, ``bench.py``'s CPU-baseline legs and ``oracle/`` scripts may import
this module. Product code under ``gaussian_process_odes_b200/`` must never import it.
"""
import torch

__version__ = "0.2.0+oracle.restatement"

_one_third = 1.0 / 3.0
_two_thirds = 2.0 / 3.0


# --------------------------------------------------------------------------------------------------------------------
# fixed-grid solvers (torchdiffeq/_impl/fixed_grid.py + rk_common.rk4_alt_step_func, release 0.2.0)
# --------------------------------------------------------------------------------------------------------------------
def _rk4_alt_step(func, t, dt, y):
    k1 = func(t, y)
    k2 = func(t + dt * _one_third, y + dt * k1 * _one_third)
    k3 = func(t + dt * _two_thirds, y + dt * (k2 - k1 * _one_third))
    k4 = func(t + dt, y + dt * (k1 - k2 + k3))
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


def _midpoint_step(func, t, dt, y):
    half_dt = 0.5 * dt
    y_mid = y + func(t, y) * half_dt
    return dt * func(t + half_dt, y_mid)


def _euler_step(func, t, dt, y):
    return dt * func(t, y)


_FIXED = {"rk4": _rk4_alt_step, "midpoint": _midpoint_step, "euler": _euler_step}


def _fixed_grid_integrate(step, func, y0, t):
    # the user-supplied grid IS the step grid (no step_size option is ever passed by the reference)
    sol = [y0]
    y = y0
    for t0, t1 in zip(t[:-1], t[1:]):
        dy = step(func, t0, t1 - t0, y)
        y = y + dy
        sol.append(y)  # linear interpolation returns y1 exactly at a grid point
    return torch.stack(sol, 0)


# --------------------------------------------------------------------------------------------------------------------
# dopri5 (torchdiffeq/_impl/dopri5.py, rk_common.py, interp.py, misc.py, release 0.2.0)
# --------------------------------------------------------------------------------------------------------------------
_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_C_ERR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1.0 / 60.0,
]
_C_MID = [
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
]


def _rms_norm(x):
    return x.pow(2).mean().sqrt()


def _rk_step(func, y0, f0, t0, dt, beta, c_err):
    """One Dormand-Prince attempt. ``dt`` arrives in float64 and is cast to the state dtype first."""
    dt = dt.to(y0.dtype)
    t0 = t0.to(y0.dtype)
    k = [f0]
    yi = y0
    for i in range(6):
        ti = t0 + _ALPHA[i] * dt
        ks = torch.stack(k, -1)  # (..., i+1)
        yi = y0 + ks.matmul(beta[i] * dt).view_as(f0)
        k.append(func(ti, yi))
    ks = torch.stack(k, -1)  # (..., 7)
    y1 = yi  # FSAL: last beta row == c_sol[:-1] and c_sol[-1] == 0
    f1 = k[-1]
    y1_error = ks.matmul(dt * c_err)
    return y1, f1, y1_error, ks


def _interp_fit(y0, y1, y_mid, f0, f1, dt):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    e = y0
    return [e, d, c, b, a]


def _interp_evaluate(coefficients, t0, t1, t):
    x = (t - t0) / (t1 - t0)
    x = x.to(coefficients[0].dtype)
    total = coefficients[0] + x * coefficients[1]
    x_power = x
    for coefficient in coefficients[2:]:
        x_power = x_power * x
        total = total + x_power * coefficient
    return total


@torch.no_grad()
def _select_initial_step(func, t0, y0, order, rtol, atol, f0):
    dtype = y0.dtype
    t_dtype = t0.dtype
    t0 = t0.to(dtype)
    scale = atol + torch.abs(y0) * rtol
    d0 = _rms_norm(y0 / scale)
    d1 = _rms_norm(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=dtype, device=y0.device)
    else:
        h0 = 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = func(t0 + h0, y1)
    d2 = _rms_norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=dtype, device=y0.device), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    return torch.min(100 * h0, h1).to(t_dtype)


@torch.no_grad()
def _optimal_step_size(last_step, error_ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
    if error_ratio == 0:
        return last_step * ifactor
    if error_ratio < 1:
        dfactor = 1.0
    error_ratio = error_ratio.type_as(last_step)
    exponent = 1.0 / order
    factor = min(ifactor, max(float(safety / error_ratio ** exponent), dfactor))
    return last_step * factor


def _dopri5_integrate(func, y0, t, rtol, atol, stats=None):
    dev, dt_y = y0.device, y0.dtype
    beta = [torch.tensor(b, dtype=torch.float64).to(dt_y).to(dev) for b in _BETA]
    c_err = torch.tensor(_C_ERR, dtype=torch.float64).to(dt_y).to(dev)
    c_mid = torch.tensor(_C_MID, dtype=torch.float64).to(dt_y).to(dev)

    t = t.to(torch.float64)  # "all time-like objects use float64"
    f0 = func(t[0], y0)
    dt = _select_initial_step(func, t[0], y0, 4, rtol, atol, f0)
    y, f, t0s, t1s = y0, f0, t[0], t[0]
    coeff = [y0] * 5
    sol = [y0]
    n_acc = n_rej = 0
    for i in range(1, len(t)):
        next_t = t[i]
        while next_t > t1s:
            assert t1s + dt > t1s, "underflow in dt {}".format(float(dt))
            t1_new = t1s + dt
            y1, f1, y1_error, ks = _rk_step(func, y, f, t1s, dt, beta, c_err)
            error_tol = atol + rtol * torch.max(y.abs(), y1.abs())
            error_ratio = _rms_norm(y1_error / error_tol).detach()
            accept = bool(error_ratio <= 1)
            if accept:
                dts = dt.to(dt_y)
                y_mid = y + ks.matmul(dts * c_mid).view_as(y)
                coeff = _interp_fit(y, y1, y_mid, ks[..., 0], ks[..., -1], dts)
                t0s, t1s = t1s, t1_new
                y, f = y1, f1
                n_acc += 1
            else:
                n_rej += 1
            dt = _optimal_step_size(dt, error_ratio)
        sol.append(_interp_evaluate(coeff, t0s, t1s, next_t))
    if stats is not None:
        stats["accepted"], stats["rejected"] = n_acc, n_rej
    return torch.stack(sol, 0)


# --------------------------------------------------------------------------------------------------------------------
# public API (torchdiffeq/_impl/odeint.py, release 0.2.0)
# --------------------------------------------------------------------------------------------------------------------
class _ReverseFunc(torch.nn.Module):
    def __init__(self, base_func):
        super().__init__()
        self.base_func = base_func

    def forward(self, t, y):
        return -self.base_func(-t, y)


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None, _stats=None):
    """Integrate ``dy/dt = func(t, y)``; returns ``(len(t), *y0.shape)`` with ``out[0] == y0``."""
    if isinstance(y0, (tuple, list)):
        raise NotImplementedError("tuple state (the reference's dead divergence branch) is not restated")
    if method is None:
        method = "dopri5"
    assert t.ndim == 1 and len(t) >= 1
    d = t[1:] - t[:-1] if len(t) > 1 else torch.ones(1)
    if len(t) > 1 and bool((d < 0).all()):
        t = -t
        func = _ReverseFunc(func)
    else:
        assert bool((d > 0).all()), "t must be strictly increasing or decreasing"
    if method in _FIXED:
        return _fixed_grid_integrate(_FIXED[method], func, y0, t)
    if method == "dopri5":
        return _dopri5_integrate(func, y0, t, rtol, atol, stats=_stats)
    raise ValueError("oracle restates only rk4 / midpoint / euler / dopri5, got %r" % (method,))


def odeint_adjoint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None, **unused):
    # never enabled by the reference's defaults (use_adjoint=False); the oracle maps it to plain autograd
    return odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
