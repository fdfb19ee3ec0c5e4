"""Transfer function, growth ODE and linear power on the host (float64 torch + autograd):
stand-in for ``pmwd/boltzmann.py:8-455`` and the Dopri5 integrator of
``pmwd/ode_util.py:103-236``.  Off the hot path (a 4-variable ODE and 1-D tables).

Differences from the reference, by necessity of having no JAX here:
  * gradients flow by torch autograd through the accepted Runge-Kutta steps
    (discretise-then-differentiate) instead of the continuous adjoint ODE of
    ``ode_util.py:238-276``; both agree to the ODE tolerance;
  * ``varlin`` / ``sigma8`` (mcfit) are not provided.
"""
import math

import torch

from .cosmology import H_deriv, Omega_m_a


# ---------------------------------------------------------------------------- transfer
def transfer_fit(k, cosmo, conf):
    """Eisenstein & Hu fit, ``pmwd/boltzmann.py:32-123``."""
    k = torch.as_tensor(k, dtype=conf.cosmo_dtype)
    k = k * cosmo.h / conf.L * conf.Mpc_SI

    T2_cmb_norm = (conf.T_cmb / 2.7) ** 2
    h2 = cosmo.h ** 2
    w_m = cosmo.Omega_m * h2
    w_b = cosmo.Omega_b * h2
    f_b = cosmo.Omega_b / cosmo.Omega_m
    f_c = cosmo.Omega_c / cosmo.Omega_m

    z_eq = 2.50e4 * w_m / T2_cmb_norm ** 2
    k_eq = 7.46e-2 * w_m / T2_cmb_norm

    b1 = 0.313 * w_m ** -0.419 * (1 + 0.607 * w_m ** 0.674)
    b2 = 0.238 * w_m ** 0.223
    z_d = 1291 * w_m ** 0.251 / (1 + 0.659 * w_m ** 0.828) * (1 + b1 * w_b ** b2)

    R_d = 31.5 * w_b / T2_cmb_norm ** 2 * (1e3 / z_d)
    R_eq = 31.5 * w_b / T2_cmb_norm ** 2 * (1e3 / z_eq)
    s = (2 / (3 * k_eq) * torch.sqrt(6 / R_eq)
         * torch.log((torch.sqrt(1 + R_d) + torch.sqrt(R_eq + R_d)) / (1 + torch.sqrt(R_eq))))
    k_silk = 1.6 * w_b ** 0.52 * w_m ** 0.73 * (1 + (10.4 * w_m) ** -0.95)

    if conf.transfer_fit_nowiggle:
        alpha_gamma = (1 - 0.328 * torch.log(431 * w_m) * f_b
                       + 0.38 * torch.log(22.3 * w_m) * f_b ** 2)
        gamma_eff_ratio = alpha_gamma + (1 - alpha_gamma) / (1 + (0.43 * k * s) ** 4)
        q_eff = k / (13.41 * k_eq * gamma_eff_ratio)
        L0 = torch.log(2 * math.e + 1.8 * q_eff)
        C0 = 14.2 + 731 / (1 + 62.5 * q_eff)
        return L0 / (L0 + C0 * q_eff ** 2)

    a1 = (46.9 * w_m) ** 0.670 * (1 + (32.1 * w_m) ** -0.532)
    a2 = (12.0 * w_m) ** 0.424 * (1 + (45.0 * w_m) ** -0.582)
    alpha_c = a1 ** -f_b * a2 ** -f_b ** 3
    b1 = 0.944 / (1 + (458 * w_m) ** -0.708)
    b2 = (0.395 * w_m) ** -0.0266
    beta_c = 1 / (1 + b1 * (f_c ** b2 - 1))

    def T0_tilde(k, alpha_c, beta_c):
        q = k / (13.41 * k_eq)
        L = torch.log(math.e + 1.8 * beta_c * q)
        C = 14.2 / alpha_c + 386 / (1 + 69.9 * q ** 1.08)
        return L / (L + C * q ** 2)

    f = 1 / (1 + (k * s / 5.4) ** 4)
    T_c = f * T0_tilde(k, 1, beta_c) + (1 - f) * T0_tilde(k, alpha_c, beta_c)

    y = (1 + z_eq) / (1 + z_d)
    x = torch.sqrt(1 + y)
    G = y * (-6 * x + (2 + 3 * y) * torch.log((x + 1) / (x - 1)))
    alpha_b = 2.07 * k_eq * s * (1 + R_d) ** -0.75 * G

    beta_node = 8.41 * w_m ** 0.435
    beta_b = 0.5 + f_b + (3 - 2 * f_b) * torch.sqrt(1 + (17.2 * w_m) ** 2)

    ks = k * s
    T_b = (T0_tilde(k, 1, 1) / (1 + (ks / 5.2) ** 2)
           + alpha_b * ks ** 3 / (beta_b ** 3 + ks ** 3) * torch.exp(-(k / k_silk) ** 1.4)
           ) * torch.sinc(ks ** 2 / (math.pi * (beta_node ** 3 + ks ** 3) ** (1 / 3)))

    return f_c * T_c + f_b * T_b


def transfer_integ(cosmo, conf):
    """``pmwd/boltzmann.py:8-29``."""
    if not conf.transfer_fit:
        raise NotImplementedError('TODO')
    return cosmo.replace(transfer=transfer_fit(conf.transfer_k, cosmo, conf))


def _interp(x, xp, fp):
    """``jnp.interp``: piecewise-linear, clamped at the ends; differentiable w.r.t. fp."""
    x = torch.as_tensor(x, dtype=fp.dtype)
    idx = torch.clamp(torch.searchsorted(xp, x.detach().contiguous(), right=True) - 1,
                      0, len(xp) - 2)
    x0, x1 = xp[idx], xp[idx + 1]
    t = torch.clamp((x - x0) / (x1 - x0), 0, 1)
    return fp[idx] + t * (fp[idx + 1] - fp[idx])


def transfer(k, cosmo, conf):
    """``pmwd/boltzmann.py:126-160``."""
    if cosmo.transfer is None:
        raise ValueError('Transfer table is empty. Call transfer_integ or boltzmann first.')
    k = torch.as_tensor(k)
    out_dtype = k.dtype if k.dtype.is_floating_point else conf.cosmo_dtype
    T = _interp(k.to(conf.cosmo_dtype).cpu(), conf.transfer_k, cosmo.transfer)
    return T.to(out_dtype)


# ---------------------------------------------------------------------------- growth
# Dopri5 tableau, pmwd/ode_util.py:103-121
_ALPHA = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.)
_BETA = (
    (1 / 5,),
    (3 / 40, 9 / 40),
    (44 / 45, -56 / 15, 32 / 9),
    (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
    (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
    (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84),
)
_C_SOL = (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0)
_C_ERR = (35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
          -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.)
_C_MID = (6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
          -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
          -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2)


def _rk_step(func, y0, f0, t0, dt):
    """``runge_kutta_step``, ``pmwd/ode_util.py:101-131``."""
    k = [f0]
    for i in range(6):
        yi = y0 + dt * sum(b * kj for b, kj in zip(_BETA[i], k) if b != 0)
        k.append(func(yi, t0 + dt * _ALPHA[i]))
    y1 = dt * sum(c * kj for c, kj in zip(_C_SOL, k) if c != 0) + y0
    y1_err = dt * sum(c * kj for c, kj in zip(_C_ERR, k) if c != 0)
    return y1, k[-1], y1_err, k


def _interp_fit(y0, y1, k, dt):
    """``interp_fit_dopri`` + ``fit_4th_order_polynomial``, ``pmwd/ode_util.py:59-76``."""
    y_mid = y0 + dt * sum(c * kj for c, kj in zip(_C_MID, k) if c != 0)
    dy0, dy1 = k[0], k[-1]
    a = -2. * dt * dy0 + 2. * dt * dy1 - 8. * y0 - 8. * y1 + 16. * y_mid
    b = 5. * dt * dy0 - 3. * dt * dy1 + 18. * y0 + 14. * y1 - 32. * y_mid
    c = -4. * dt * dy0 + dt * dy1 - 11. * y0 - 5. * y1 + 16. * y_mid
    d = dt * dy0
    return a, b, c, d, y0


def _initial_step_size(func, t0, y0, order, rtol, atol, f0):
    """``initial_step_size``, ``pmwd/ode_util.py:78-99`` (value only, no gradient)."""
    with torch.no_grad():
        scale = atol + y0.abs() * rtol
        d0 = torch.linalg.norm(y0 / scale).item()
        d1 = torch.linalg.norm(f0 / scale).item()
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        f1 = func(y0 + h0 * f0, t0 + h0)
        d2 = torch.linalg.norm((f1 - f0) / scale).item() / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1. / (order + 1.))
    return min(100. * h0, h1)


def odeint(func, y0, ts, rtol, atol, dt0=None, mxstep=100000):
    """Adaptive Dopri5 with dense output at ``ts``: the forward algorithm of
    ``pmwd/ode_util.py:173-212`` (same tableau, error norm, step controller and 4th-order
    interpolant).  Step-size control is not differentiated; the state is."""
    t = float(ts[0])
    y = y0
    f = func(y, t)
    dt = dt0[0] if isinstance(dt0, tuple) else dt0
    if dt is None:
        dt = _initial_step_size(func, t, y, 4, rtol, atol, f)
    dt = max(float(dt), 0.)
    last_t = t
    coeff = (y, y, y, y, y)
    outs = [y0]
    for target in ts[1:]:
        target = float(target)
        n = 0
        while t < target and n < mxstep and dt > 0:
            y1, f1, y1_err, k = _rk_step(func, y, f, t, dt)
            with torch.no_grad():   # mean_error_ratio, ode_util.py:139-142
                tol = atol + rtol * torch.maximum(y.abs(), y1.abs())
                ratio = torch.sqrt(torch.mean((y1_err / tol) ** 2)).item()
            if ratio <= 1.:
                coeff = _interp_fit(y, y1, k, dt)
                last_t, t, y, f = t, t + dt, y1, f1
            # optimal_step_size, ode_util.py:144-152
            dfactor = 1.0 if ratio < 1 else 0.2
            if ratio == 0:
                dt = dt * 10.0
            else:
                dt = dt * min(10.0, max(ratio ** (-1.0 / 5.0) * 0.9, dfactor))
            n += 1
        rel = (target - last_t) / (t - last_t) if t != last_t else 0.
        a, b, c, d, e = coeff
        outs.append((((a * rel + b) * rel + c) * rel + d) * rel + e)   # jnp.polyval
    return torch.stack(outs)


def growth_integ(cosmo, conf):
    """``pmwd/boltzmann.py:163-229``: growth table of shape (2, 3, len(conf.growth_a))."""
    dtype = conf.cosmo_dtype
    eps = torch.finfo(dtype).eps
    a_ic = 0.5 * eps ** (1 / 3)
    if a_ic >= conf.a_lpt_step:
        a_ic = 0.1 * conf.a_lpt_step

    a = conf.growth_a.clone()
    a[0] = a_ic
    lna = torch.log(a)
    num_order, num_deriv, num_a = 2, 3, len(a)

    def ode(G, lna_):
        a_ = torch.exp(torch.as_tensor(lna_, dtype=dtype))
        dlnH_dlna = H_deriv(a_, cosmo)
        Omega_fac = 1.5 * Omega_m_a(a_, cosmo)
        G1, G1p, G2, G2p = G.unbind(-1)
        G1pp = -(3 + dlnH_dlna - Omega_fac) * G1 - (4 + dlnH_dlna) * G1p
        G2pp = Omega_fac * G1 ** 2 - (8 + 2 * dlnH_dlna - Omega_fac) * G2 - (6 + dlnH_dlna) * G2p
        return torch.stack((G1p, G1pp, G2p, G2pp), dim=-1)

    G_ic = torch.tensor((1, 0, 3 / 7, 0), dtype=dtype)
    G = odeint(ode, G_ic, lna.tolist(), conf.growth_rtol, conf.growth_atol,
               dt0=conf.growth_inistep)
    G_deriv = ode(G, lna)

    G = G.reshape(num_a, num_order, num_deriv - 1)
    G_deriv = G_deriv.reshape(num_a, num_order, num_deriv - 1)
    G = torch.cat((G, G_deriv[..., -1:]), dim=2)
    G = G.movedim(0, 2)

    m = torch.tensor((1., 2.), dtype=dtype)[:, None]
    growth = torch.stack((
        G[:, 0],
        m * G[:, 0] + G[:, 1],
        m ** 2 * G[:, 0] + 2 * m * G[:, 1] + G[:, 2],
    ), dim=1)
    return cosmo.replace(growth=growth)


def growth(a, cosmo, conf, order=1, deriv=0):
    """``pmwd/boltzmann.py:233-269``: ``a**order * interp(a, conf.growth_a, table)``."""
    if cosmo.growth is None:
        raise ValueError('Growth table is empty. Call growth_integ or boltzmann first.')
    a = torch.as_tensor(a, dtype=conf.cosmo_dtype)
    return a ** order * _interp(a, conf.growth_a, cosmo.growth[order - 1][deriv])


def boltzmann(cosmo, conf, transfer=True, growth=True, varlin=False):
    """``pmwd/boltzmann.py:338-374``.  ``varlin`` (mcfit) is not provided."""
    if varlin:
        raise NotImplementedError('varlin needs mcfit, which is off the hot path and not '
                                  'available in this image')
    cosmo = transfer_integ(cosmo, conf) if transfer else cosmo.replace(transfer=None)
    cosmo = growth_integ(cosmo, conf) if growth else cosmo.replace(growth=None)
    return cosmo.replace(varlin=None)


# ---------------------------------------------------------------------------- power
class _SafePower(torch.autograd.Function):
    """``_safe_power``, ``pmwd/boltzmann.py:377-396``: x1**x2 with finite gradients at 0."""

    @staticmethod
    def forward(ctx, x1, x2):
        y = x1 ** x2
        ctx.save_for_backward(x1, x2, y)
        return y

    @staticmethod
    def backward(ctx, y_cot):
        x1, x2, y = ctx.saved_tensors
        nz = x1 != 0
        safe = torch.where(nz, x1, torch.ones_like(x1))
        x1_cot = torch.where(nz, x2 * y / safe * y_cot, torch.zeros_like(y))
        lnx1 = torch.where(nz, torch.log(safe), torch.zeros_like(x1))
        x2_cot = (lnx1 * y * y_cot).sum()
        return x1_cot, x2_cot


def linear_power(k, a, cosmo, conf):
    """``pmwd/boltzmann.py:399-455``.  Evaluated where ``k`` lives (CPU or CUDA) in
    float64, returned in ``k``'s float dtype."""
    if conf.dim != 3:
        raise ValueError(f'dim={conf.dim} not supported')
    k = torch.as_tensor(k)
    out_dtype = k.dtype if k.dtype.is_floating_point else conf.cosmo_dtype
    dev = k.device
    k64 = k.to(conf.cosmo_dtype)

    if cosmo.transfer is None:
        raise ValueError('Transfer table is empty. Call transfer_integ or boltzmann first.')
    T = _interp(k64, conf.transfer_k.to(dev), cosmo.transfer.to(dev)).to(out_dtype).to(conf.cosmo_dtype)

    k_pivot = cosmo.k_pivot.to(dev)
    Plin = (0.32 * cosmo.A_s.to(dev) * k_pivot
            * _SafePower.apply(k64 / k_pivot, cosmo.n_s.to(dev))
            * (math.pi * (conf.c / conf.H_0) ** 2 / cosmo.Omega_m.to(dev) * T) ** 2)
    if a is not None:
        D = growth(a, cosmo, conf).to(dev)
        Plin = Plin * D ** 2
    return Plin.to(out_dtype)
