"""Oracle initial conditions: restates ``pmwd/modes.py:12-117`` and
``pmwd/lpt.py:13-76,136-212``.  TEST INFRASTRUCTURE ONLY.

``white_noise`` cannot reproduce ``jax.random.normal(PRNGKey(seed))``
(``pmwd/modes.py:33-36``) without JAX; as SURVEY.md 8(d) prescribes the synthetic
stream is ``numpy.random.default_rng(seed).standard_normal(shape, float32)``.
"""
import numpy as np

from . import cosmo as oc
from . import pm
from .gravity import fftfreq, fftfwd, fftinv, laplace, neg_grad


def white_noise(seed, conf, real=False):
    """``pmwd/modes.py:12-49`` (``unit_abs=False``)."""
    rng = np.random.default_rng(seed)
    modes = rng.standard_normal(conf.ptcl_grid_shape, dtype=np.float32)
    modes = modes.astype(conf.float_dtype)
    if real:
        return modes
    return fftfwd(modes, norm='ortho')


def linear_modes(modes, cosmo, conf, a=None, real=False):
    """``pmwd/modes.py:67-117``: ``modes * sqrt(P_lin(k) * V)``."""
    fdt = conf.float_dtype
    kvec = fftfreq(conf.ptcl_grid_shape, conf.ptcl_spacing, dtype=fdt)
    k = np.sqrt(sum(k ** 2 for k in kvec))
    if a is not None:
        a = fdt.type(a)
    Plin = oc.linear_power(k, a, cosmo, conf)
    if np.isrealobj(modes):
        modes = fftfwd(modes, norm='ortho')
    modes = modes * np.sqrt(Plin * fdt.type(conf.box_vol)).astype(fdt)
    if real:
        modes = fftinv(modes, shape=conf.ptcl_grid_shape, norm=conf.ptcl_spacing)
    return modes


def _strain(kvec, i, j, pot, conf):
    """``pmwd/lpt.py:13-37``."""
    k_i, k_j = kvec[i], kvec[j]
    nyquist = np.pi / conf.ptcl_spacing
    eps = nyquist * np.finfo(conf.float_dtype).eps
    if i != j:
        k_i = np.where(np.abs(np.abs(k_i) - nyquist) <= eps, 0, k_i)
        k_j = np.where(np.abs(np.abs(k_j) - nyquist) <= eps, 0, k_j)
    strain = (-k_i * k_j * pot).astype(pot.dtype)
    strain = fftinv(strain, shape=conf.ptcl_grid_shape)
    return strain.astype(conf.float_dtype)


def _L(kvec, pot_m, pot_n, conf):
    """``pmwd/lpt.py:40-76`` (only the ``m == n`` case is reachable for 2LPT)."""
    m_eq_n = pot_n is None
    if m_eq_n:
        pot_n = pot_m
    L = np.zeros(conf.ptcl_grid_shape, dtype=conf.float_dtype)
    for i in range(conf.dim):
        strain_m = _strain(kvec, i, i, pot_m, conf)
        for j in range(conf.dim - 1, i, -1):
            strain_n = _strain(kvec, j, j, pot_n, conf)
            L = L + strain_m * strain_n
        if not m_eq_n:
            for j in range(i - 1, -1, -1):
                strain_n = _strain(kvec, j, j, pot_n, conf)
                L = L + strain_m * strain_n
    if not m_eq_n:
        L = L * 0.5
    for i in range(conf.dim - 1):
        for j in range(i + 1, conf.dim):
            strain_m = _strain(kvec, i, j, pot_m, conf)
            strain_n = strain_m
            if not m_eq_n:
                strain_n = _strain(kvec, j, i, pot_n, conf)
            L = L - strain_m * strain_n
    return L


def lpt(modes, cosmo, conf):
    """``pmwd/lpt.py:136-212``: returns ``dict(pmid, disp, vel)``."""
    if conf.dim not in (1, 2, 3):
        raise ValueError(f'dim={conf.dim} not supported')
    if conf.lpt_order not in (0, 1, 2, 3):
        raise ValueError(f'lpt_order={conf.lpt_order} not supported')
    fdt = conf.float_dtype
    cdt = np.result_type(fdt, np.complex64)
    modes = (modes / fdt.type(conf.ptcl_cell_vol)).astype(cdt)        # lpt.py:164
    kvec = fftfreq(conf.ptcl_grid_shape, conf.ptcl_spacing, dtype=fdt)

    pot = []
    if conf.lpt_order > 0:
        pot_1 = laplace(kvec, modes)                                  # :173
        pot.append(pot_1)
    if conf.lpt_order > 1:
        src_2 = _L(kvec, pot_1, None, conf)                           # :177
        src_2 = fftfwd(src_2)                                         # :179
        pot.append(laplace(kvec, src_2))                              # :181
    if conf.lpt_order > 2:
        raise NotImplementedError('TODO')

    a = conf.a_start
    pmid, disp, vel, _ = pm.gen_grid(conf, vel=True)                  # :188
    for order in range(1, 1 + conf.lpt_order):
        D = oc.growth(a, cosmo, conf, order=order)
        dD_dlna = oc.growth(a, cosmo, conf, order=order, deriv=1)
        a2HDp = a ** 2 * np.sqrt(oc.E2(a, cosmo)) * dD_dlna
        D = fdt.type(D)
        a2HDp = fdt.type(a2HDp)
        for i, k in enumerate(kvec):
            grad = neg_grad(k, pot[order - 1], conf.ptcl_spacing)     # :199
            grad = fftinv(grad, shape=conf.ptcl_grid_shape).astype(fdt)
            grad = grad.ravel()
            disp[:, i] = disp[:, i] + D * grad                        # :206
            vel[:, i] = vel[:, i] + a2HDp * grad                      # :207
    return dict(pmid=pmid, disp=disp, vel=vel)


# ----------------------------------------------------------------------------
# VJP of  real white noise -> linear_modes -> lpt  w.r.t. the white noise
# ----------------------------------------------------------------------------

def _circ(field, mult, conf):
    """Real circulant operator: irfftn(mult * rfftn(field))."""
    return fftinv(mult * fftfwd(field), shape=conf.ptcl_grid_shape)


def lpt_vjp_modes(white, cosmo, conf, disp_cot, vel_cot):
    """Cotangent of a *real* white-noise field (``white_noise(seed, conf, real=True)``)
    through ``linear_modes`` + ``lpt`` (what ``jax.grad`` computes by AD through
    ``pmwd/modes.py:67-84`` and ``pmwd/lpt.py:136-212``), written with the
    transposes of the real circulant operators involved: a multiplier m(k) has
    transpose conj(m(k)), so the -ik gradient flips sign and the k_i k_j / k^2
    operators are self-adjoint.  Float64 recommended.  Checked against finite
    differences in ``tests/test_oracle_vjp.py``.
    """
    fdt = conf.float_dtype
    shape = conf.ptcl_grid_shape
    N = conf.ptcl_num
    kvec = fftfreq(shape, conf.ptcl_spacing, dtype=fdt)
    k2 = sum(k ** 2 for k in kvec)
    kk = np.sqrt(k2)
    Plin = oc.linear_power(kk, None, cosmo, conf)
    amp = np.sqrt(Plin * conf.box_vol) / conf.ptcl_cell_vol / np.sqrt(N)
    with np.errstate(divide='ignore', invalid='ignore'):
        inv_lap = np.where(k2 != 0, -1 / k2, 0)
    m_phi1 = amp * inv_lap                       # white -> phi1 (real, symmetric)

    nyq = np.pi / conf.ptcl_spacing
    eps = nyq * np.finfo(fdt).eps
    kmask = [np.where(np.abs(np.abs(k) - nyq) <= eps, 0, k) for k in kvec]

    def grad_T(g, i):       # transpose of circ(-i k_i masked) = circ(+i k_i masked)
        return _circ(g, 1j * kmask[i], conf)

    def strain_mult(i, j):
        if i == j:
            return -kvec[i] * kvec[j]
        return -kmask[i] * kmask[j]

    a = conf.a_start
    fac = []
    for order in (1, 2):
        D = oc.growth(a, cosmo, conf, order=order)
        a2HDp = a ** 2 * np.sqrt(oc.E2(a, cosmo)) * oc.growth(a, cosmo, conf, order=order, deriv=1)
        fac.append((D, a2HDp))

    dc = np.asarray(disp_cot, dtype=fdt).reshape(shape + (3,))
    vc = np.asarray(vel_cot, dtype=fdt).reshape(shape + (3,))

    phi1 = _circ(white.astype(fdt), m_phi1, conf)
    phi1_cot = np.zeros(shape, dtype=fdt)
    for i in range(3):
        phi1_cot += grad_T(fac[0][0] * dc[..., i] + fac[0][1] * vc[..., i], i)

    if conf.lpt_order > 1:
        phi2_cot = np.zeros(shape, dtype=fdt)
        for i in range(3):
            phi2_cot += grad_T(fac[1][0] * dc[..., i] + fac[1][1] * vc[..., i], i)
        L_cot = _circ(phi2_cot, inv_lap, conf)
        s = {}
        for i in range(3):
            for j in range(i, 3):
                s[i, j] = _circ(phi1, strain_mult(i, j), conf)
        for i in range(3):
            others = sum(s[j, j] for j in range(3) if j != i)
            phi1_cot += _circ(L_cot * others, strain_mult(i, i), conf)
        for i in range(2):
            for j in range(i + 1, 3):
                phi1_cot += _circ(-2 * L_cot * s[i, j], strain_mult(i, j), conf)

    return _circ(phi1_cot, m_phi1, conf)
