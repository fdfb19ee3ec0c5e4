"""Oracle host-side cosmology in float64: restates ``pmwd/cosmology.py:185-269``
and ``pmwd/boltzmann.py:32-123,163-269,399-455``.  TEST INFRASTRUCTURE ONLY.

The reference integrates the growth ODE with its own Dopri5 fork
(``pmwd/ode_util.py``) at rtol = atol = sqrt(eps_f64); here ``scipy`` DOP853 at
1e-11 is used, so tables agree to ~1e-8 relative (the reference's tolerance).
"""
import copy

import numpy as np
from scipy.integrate import solve_ivp


class Cosmo:
    """Parameters of ``pmwd/cosmology.py:53-72`` (float64)."""

    def __init__(self, conf, A_s_1e9, n_s, Omega_m, Omega_b, h,
                 Omega_k_=None, w_0_=None, w_a_=None):
        self.conf = conf
        self.A_s_1e9 = float(A_s_1e9)
        self.n_s = float(n_s)
        self.Omega_m = float(Omega_m)
        self.Omega_b = float(Omega_b)
        self.h = float(h)
        self.Omega_k_ = Omega_k_
        self.w_0_ = w_0_
        self.w_a_ = w_a_
        self.transfer = None
        self.growth = None

    def replace(self, **kw):
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new

    # cosmology.py:112-150
    @property
    def k_pivot(self):
        return self.conf.k_pivot_Mpc / (self.h * self.conf.Mpc_SI) * self.conf.L

    @property
    def A_s(self):
        return self.A_s_1e9 * 1e-9

    @property
    def Omega_c(self):
        return self.Omega_m - self.Omega_b

    @property
    def Omega_k(self):
        return 0. if self.Omega_k_ is None else self.Omega_k_

    @property
    def Omega_de(self):
        return 1 - (self.Omega_m + self.Omega_k)

    @property
    def w_0(self):
        return -1. if self.w_0_ is None else self.w_0_

    @property
    def w_a(self):
        return 0. if self.w_a_ is None else self.w_a_


def SimpleLCDM(conf, **kw):
    """``pmwd/cosmology.py:164-171``."""
    p = dict(A_s_1e9=2.0, n_s=0.96, Omega_m=0.3, Omega_b=0.05, h=0.7)
    p.update(kw)
    return Cosmo(conf, **p)


def E2(a, cosmo):
    """``pmwd/cosmology.py:185-219``."""
    a = np.asarray(a, dtype=np.float64)
    de_a = a ** (-3 * (1 + cosmo.w_0 + cosmo.w_a)) * np.exp(-3 * cosmo.w_a * (1 - a))
    return cosmo.Omega_m * a ** -3 + cosmo.Omega_k * a ** -2 + cosmo.Omega_de * de_a


def H_deriv(a, cosmo):
    """dlnH/dlna (``pmwd/cosmology.py:222-242``); the reference differentiates
    ``E2`` with JAX AD, here the derivative is written out."""
    a = np.asarray(a, dtype=np.float64)
    p = -3 * (1 + cosmo.w_0 + cosmo.w_a)
    de_a = a ** p * np.exp(-3 * cosmo.w_a * (1 - a))
    dde = de_a * (p / a + 3 * cosmo.w_a)
    dE2 = -3 * cosmo.Omega_m * a ** -4 - 2 * cosmo.Omega_k * a ** -3 + cosmo.Omega_de * dde
    return 0.5 * a * dE2 / E2(a, cosmo)


def Omega_m_a(a, cosmo):
    """``pmwd/cosmology.py:245-269``."""
    a = np.asarray(a, dtype=np.float64)
    return cosmo.Omega_m / (a ** 3 * E2(a, cosmo))


def growth_integ(cosmo, conf):
    """``pmwd/boltzmann.py:163-229``: table of shape (2 orders, 3 derivs, len(growth_a))."""
    eps = np.finfo(np.float64).eps
    a_ic = 0.5 * np.cbrt(eps)
    if a_ic >= conf.a_lpt_step:
        a_ic = 0.1 * conf.a_lpt_step
    a = conf.growth_a.copy()
    a[0] = a_ic
    lna = np.log(a)

    def ode(lna_, G):
        a_ = np.exp(lna_)
        dlnH = H_deriv(a_, cosmo)
        Of = 1.5 * Omega_m_a(a_, cosmo)
        G1, G1p, G2, G2p = G
        G1pp = -(3 + dlnH - Of) * G1 - (4 + dlnH) * G1p
        G2pp = Of * G1 ** 2 - (8 + 2 * dlnH - Of) * G2 - (6 + dlnH) * G2p
        return np.array([G1p, G1pp, G2p, G2pp])

    G_ic = np.array([1, 0, 3 / 7, 0], dtype=np.float64)
    sol = solve_ivp(ode, (lna[0], lna[-1]), G_ic, method='DOP853', t_eval=lna,
                    rtol=1e-11, atol=1e-11)
    G = sol.y.T                                    # (num_a, 4)
    Gd = np.stack([ode(l, g) for l, g in zip(lna, G)])
    num_a = len(a)
    G = G.reshape(num_a, 2, 2)
    Gd = Gd.reshape(num_a, 2, 2)
    G = np.concatenate((G, Gd[..., -1:]), axis=2)  # (num_a, order, deriv)
    G = np.moveaxis(G, 0, 2)                       # (order, deriv, num_a)
    m = np.array((1., 2.))[:, np.newaxis]
    table = np.stack((
        G[:, 0],
        m * G[:, 0] + G[:, 1],
        m ** 2 * G[:, 0] + 2 * m * G[:, 1] + G[:, 2],
    ), axis=1)
    return cosmo.replace(growth=table)


def growth(a, cosmo, conf, order=1, deriv=0):
    """``pmwd/boltzmann.py:233-269``: ``a**order * interp(a, growth_a, table)``."""
    if cosmo.growth is None:
        raise ValueError('Growth table is empty. Call growth_integ or boltzmann first.')
    a = np.asarray(a, dtype=np.float64)
    return a ** order * np.interp(a, conf.growth_a, cosmo.growth[order - 1][deriv])


def transfer_fit(k, cosmo, conf):
    """Eisenstein & Hu fit (``pmwd/boltzmann.py:32-123``)."""
    k = np.asarray(k, dtype=np.float64)
    k = k * cosmo.h / conf.L * conf.Mpc_SI

    T2 = (conf.T_cmb / 2.7) ** 2
    h2 = cosmo.h ** 2
    w_m = cosmo.Omega_m * h2
    w_b = cosmo.Omega_b * h2
    f_b = cosmo.Omega_b / cosmo.Omega_m
    f_c = cosmo.Omega_c / cosmo.Omega_m

    z_eq = 2.50e4 * w_m / T2 ** 2
    k_eq = 7.46e-2 * w_m / T2

    b1 = 0.313 * w_m ** -0.419 * (1 + 0.607 * w_m ** 0.674)
    b2 = 0.238 * w_m ** 0.223
    z_d = 1291 * w_m ** 0.251 / (1 + 0.659 * w_m ** 0.828) * (1 + b1 * w_b ** b2)

    R_d = 31.5 * w_b / T2 ** 2 * (1e3 / z_d)
    R_eq = 31.5 * w_b / T2 ** 2 * (1e3 / z_eq)
    s = (2 / (3 * k_eq) * np.sqrt(6 / R_eq)
         * np.log((np.sqrt(1 + R_d) + np.sqrt(R_eq + R_d)) / (1 + np.sqrt(R_eq))))
    k_silk = 1.6 * w_b ** 0.52 * w_m ** 0.73 * (1 + (10.4 * w_m) ** -0.95)

    if conf.transfer_fit_nowiggle:
        alpha_gamma = (1 - 0.328 * np.log(431 * w_m) * f_b
                       + 0.38 * np.log(22.3 * w_m) * f_b ** 2)
        gamma_eff_ratio = alpha_gamma + (1 - alpha_gamma) / (1 + (0.43 * k * s) ** 4)
        q_eff = k / (13.41 * k_eq * gamma_eff_ratio)
        L0 = np.log(2 * np.e + 1.8 * q_eff)
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
        L = np.log(np.e + 1.8 * beta_c * q)
        C = 14.2 / alpha_c + 386 / (1 + 69.9 * q ** 1.08)
        return L / (L + C * q ** 2)

    f = 1 / (1 + (k * s / 5.4) ** 4)
    T_c = f * T0_tilde(k, 1, beta_c) + (1 - f) * T0_tilde(k, alpha_c, beta_c)

    y = (1 + z_eq) / (1 + z_d)
    x = np.sqrt(1 + y)
    G = y * (-6 * x + (2 + 3 * y) * np.log((x + 1) / (x - 1)))
    alpha_b = 2.07 * k_eq * s * (1 + R_d) ** -0.75 * G

    beta_node = 8.41 * w_m ** 0.435
    beta_b = 0.5 + f_b + (3 - 2 * f_b) * np.sqrt(1 + (17.2 * w_m) ** 2)

    T_b = (T0_tilde(k, 1, 1) / (1 + (k * s / 5.2) ** 2)
           + alpha_b * (k * s) ** 3 / (beta_b ** 3 + (k * s) ** 3)
           * np.exp(-(k / k_silk) ** 1.4)
           ) * np.sinc((k * s) ** 2 / (np.pi * np.cbrt(beta_node ** 3 + (k * s) ** 3)))

    return f_c * T_c + f_b * T_b


def transfer_integ(cosmo, conf):
    """``pmwd/boltzmann.py:8-29``."""
    return cosmo.replace(transfer=transfer_fit(conf.transfer_k, cosmo, conf))


def boltzmann(cosmo, conf):
    """``pmwd/boltzmann.py:338-374`` with ``varlin=False`` (sigma8 / mcfit is off
    the hot path)."""
    return growth_integ(transfer_integ(cosmo, conf), conf)


def linear_power(k, a, cosmo, conf):
    """``pmwd/boltzmann.py:399-455``: transfer interpolated linearly in k over
    ``conf.transfer_k`` (``:126-160``)."""
    k = np.asarray(k)
    fdt = np.promote_types(k.dtype, np.float64) if k.dtype.kind != 'f' else k.dtype
    # boltzmann.py:126-160: interpolated in float64, then cast to k's float dtype
    T = np.interp(k.astype(np.float64), conf.transfer_k, cosmo.transfer).astype(fdt)
    kk = k.astype(np.float64)
    Plin = (0.32 * cosmo.A_s * cosmo.k_pivot * (kk / cosmo.k_pivot) ** cosmo.n_s
            * (np.pi * (conf.c / conf.H_0) ** 2 / cosmo.Omega_m * T) ** 2)
    if a is not None:
        Plin = Plin * growth(a, cosmo, conf) ** 2
    return Plin.astype(fdt)
