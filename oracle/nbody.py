"""Oracle N-body integrator and its reverse-time adjoint: restates
``pmwd/nbody.py:12-162,193-276``.  TEST INFRASTRUCTURE ONLY.

Particles are plain tuples/dicts of NumPy arrays: ``pmid (N,3) int16`` and
``disp / vel / acc (N,3)`` in ``conf.float_dtype``.  The cosmology cotangent is a
dict with the two leaves that ``nbody`` can touch for a flat LCDM ``Cosmology``:
``'Omega_m'`` (float64 scalar) and ``'growth'`` (table-shaped float64 array);
every other leaf of the reference's cotangent pytree is identically zero.
"""
import numpy as np

from . import cosmo as oc
from .gravity import gravity, gravity_vjp


def G_D(a, cosmo, conf):
    """``pmwd/nbody.py:12-14``."""
    return a ** 2 * np.sqrt(oc.E2(a, cosmo)) * oc.growth(a, cosmo, conf, deriv=1)


def G_K(a, cosmo, conf):
    """``pmwd/nbody.py:17-22``."""
    return a ** 3 * oc.E2(a, cosmo) * (
        oc.growth(a, cosmo, conf, deriv=2)
        + (2 + oc.H_deriv(a, cosmo)) * oc.growth(a, cosmo, conf, deriv=1))


def drift_factor(a_vel, a_prev, a_next, cosmo, conf):
    """``pmwd/nbody.py:25-29``."""
    factor = oc.growth(a_next, cosmo, conf) - oc.growth(a_prev, cosmo, conf)
    return factor / G_D(a_vel, cosmo, conf)


def kick_factor(a_acc, a_prev, a_next, cosmo, conf):
    """``pmwd/nbody.py:32-36``."""
    factor = G_D(a_next, cosmo, conf) - G_D(a_prev, cosmo, conf)
    return factor / G_K(a_acc, cosmo, conf)


def factor_grads(fun, a0, a1, a2, cosmo, conf):
    """Value and gradient of a step factor w.r.t. the cosmology leaves it depends
    on -- what ``value_and_grad(..., argnums=3)`` returns at ``pmwd/nbody.py:51-52,
    82-83`` -- by float64 central differences (the factors are smooth rational
    functions of ``Omega_m`` and linear-rational in the growth-table entries).

    Returns ``(value, {'Omega_m': d, 'growth': table-shaped array})``.
    """
    val = fun(a0, a1, a2, cosmo, conf)
    h = 1e-6
    cp = cosmo.replace(Omega_m=cosmo.Omega_m + h)
    cm = cosmo.replace(Omega_m=cosmo.Omega_m - h)
    dOm = (fun(a0, a1, a2, cp, conf) - fun(a0, a1, a2, cm, conf)) / (2 * h)

    table = cosmo.growth
    dtab = np.zeros_like(table)
    xa = conf.growth_a
    nodes = set()
    for a in (a0, a1, a2):
        j = int(np.clip(np.searchsorted(xa, a, side='right') - 1, 0, len(xa) - 2))
        nodes.update((j, j + 1))
    for j in sorted(nodes):
        for d in range(3):
            t0 = table[0, d, j]
            step = 1e-6 * max(abs(t0), 1e-3)
            tp = table.copy(); tp[0, d, j] = t0 + step
            tm = table.copy(); tm[0, d, j] = t0 - step
            fp = fun(a0, a1, a2, cosmo.replace(growth=tp), conf)
            fm = fun(a0, a1, a2, cosmo.replace(growth=tm), conf)
            dtab[0, d, j] = (fp - fm) / (2 * step)
    return val, {'Omega_m': dOm, 'growth': dtab}


def drift(a_vel, a_prev, a_next, ptcl, cosmo, conf):
    """``pmwd/nbody.py:39-46``."""
    factor = conf.float_dtype.type(drift_factor(a_vel, a_prev, a_next, cosmo, conf))
    ptcl = dict(ptcl)
    ptcl['disp'] = ptcl['disp'] + ptcl['vel'] * factor
    return ptcl


def kick(a_acc, a_prev, a_next, ptcl, cosmo, conf):
    """``pmwd/nbody.py:70-77``."""
    factor = conf.float_dtype.type(kick_factor(a_acc, a_prev, a_next, cosmo, conf))
    ptcl = dict(ptcl)
    ptcl['vel'] = ptcl['vel'] + ptcl['acc'] * factor
    return ptcl


def _cot_axpy(cot, alpha, grads):
    """cot -= alpha * grads (leafwise, float64)."""
    return {k: cot[k] - alpha * grads[k] for k in cot}


def drift_adj(a_vel, a_prev, a_next, ptcl, ptcl_cot, cosmo, cosmo_cot, conf):
    """``pmwd/nbody.py:49-67``."""
    factor, grads = factor_grads(drift_factor, a_vel, a_prev, a_next, cosmo, conf)
    factor = conf.float_dtype.type(factor)
    ptcl = dict(ptcl)
    ptcl['disp'] = ptcl['disp'] + ptcl['vel'] * factor
    ptcl_cot = dict(ptcl_cot)
    ptcl_cot['vel'] = ptcl_cot['vel'] - ptcl_cot['disp'] * factor
    s = np.float64((ptcl_cot['disp'] * ptcl['vel']).sum())
    cosmo_cot = _cot_axpy(cosmo_cot, s, grads)
    return ptcl, ptcl_cot, cosmo_cot


def kick_adj(a_acc, a_prev, a_next, ptcl, ptcl_cot, cosmo, cosmo_cot, cosmo_cot_force, conf):
    """``pmwd/nbody.py:80-99``."""
    factor, grads = factor_grads(kick_factor, a_acc, a_prev, a_next, cosmo, conf)
    factor = conf.float_dtype.type(factor)
    ptcl = dict(ptcl)
    ptcl['vel'] = ptcl['vel'] + ptcl['acc'] * factor
    ptcl_cot = dict(ptcl_cot)
    ptcl_cot['disp'] = ptcl_cot['disp'] - ptcl_cot['acc'] * factor
    s = np.float64((ptcl_cot['vel'] * ptcl['acc']).sum())
    cosmo_cot = _cot_axpy(cosmo_cot, s, grads)
    cosmo_cot = dict(cosmo_cot)
    cosmo_cot['Omega_m'] = cosmo_cot['Omega_m'] - cosmo_cot_force * np.float64(factor)
    return ptcl, ptcl_cot, cosmo_cot


def force(a, ptcl, cosmo, conf):
    """``pmwd/nbody.py:102-105``."""
    ptcl = dict(ptcl)
    ptcl['acc'] = gravity(ptcl['pmid'], ptcl['disp'], cosmo.Omega_m, conf)
    return ptcl


def force_adj(a, ptcl, ptcl_cot, cosmo, conf):
    """``pmwd/nbody.py:108-118``."""
    acc, disp_cot, Om_cot = gravity_vjp(ptcl['pmid'], ptcl['disp'], cosmo.Omega_m, conf,
                                        ptcl_cot['vel'])
    ptcl = dict(ptcl); ptcl['acc'] = acc
    ptcl_cot = dict(ptcl_cot); ptcl_cot['acc'] = disp_cot
    return ptcl, ptcl_cot, Om_cot


def integrate(a_prev, a_next, ptcl, cosmo, conf):
    """``pmwd/nbody.py:121-140``."""
    D = K = 0
    a_disp = a_vel = a_acc = a_prev
    for d, k in conf.symp_splits:
        if d != 0:
            D += d
            a_disp_next = a_prev * (1 - D) + a_next * D
            ptcl = drift(a_vel, a_disp, a_disp_next, ptcl, cosmo, conf)
            a_disp = a_disp_next
            ptcl = force(a_disp, ptcl, cosmo, conf)
            a_acc = a_disp
        if k != 0:
            K += k
            a_vel_next = a_prev * (1 - K) + a_next * K
            ptcl = kick(a_acc, a_vel, a_vel_next, ptcl, cosmo, conf)
            a_vel = a_vel_next
    return ptcl


def integrate_adj(a_prev, a_next, ptcl, ptcl_cot, cosmo, cosmo_cot, cosmo_cot_force, conf):
    """``pmwd/nbody.py:143-162``."""
    K = D = 0
    a_disp = a_vel = a_acc = a_prev
    for d, k in reversed(conf.symp_splits):
        if k != 0:
            K += k
            a_vel_next = a_prev * (1 - K) + a_next * K
            ptcl, ptcl_cot, cosmo_cot = kick_adj(a_acc, a_vel, a_vel_next, ptcl, ptcl_cot,
                                                 cosmo, cosmo_cot, cosmo_cot_force, conf)
            a_vel = a_vel_next
        if d != 0:
            D += d
            a_disp_next = a_prev * (1 - D) + a_next * D
            ptcl, ptcl_cot, cosmo_cot = drift_adj(a_vel, a_disp, a_disp_next, ptcl,
                                                  ptcl_cot, cosmo, cosmo_cot, conf)
            a_disp = a_disp_next
            ptcl, ptcl_cot, cosmo_cot_force = force_adj(a_disp, ptcl, ptcl_cot, cosmo, conf)
            a_acc = a_disp
    return ptcl, ptcl_cot, cosmo_cot, cosmo_cot_force


def nbody_init(a, ptcl, cosmo, conf):
    """``pmwd/nbody.py:193-201``."""
    return force(a, ptcl, cosmo, conf)


def nbody_step(a_prev, a_next, ptcl, cosmo, conf):
    """``pmwd/nbody.py:204-212``."""
    return integrate(a_prev, a_next, ptcl, cosmo, conf)


def nbody(ptcl, cosmo, conf, reverse=False):
    """``pmwd/nbody.py:215-223``."""
    a_nbody = conf.a_nbody[::-1] if reverse else conf.a_nbody
    ptcl = nbody_init(a_nbody[0], ptcl, cosmo, conf)
    for a_prev, a_next in zip(a_nbody[:-1], a_nbody[1:]):
        ptcl = nbody_step(a_prev, a_next, ptcl, cosmo, conf)
    return ptcl


def nbody_adj(ptcl, ptcl_cot, cosmo, conf, reverse=False):
    """``pmwd/nbody.py:226-260``: returns ``(ptcl, ptcl_cot, cosmo_cot)``."""
    a_nbody = conf.a_nbody[::-1] if reverse else conf.a_nbody
    ptcl, ptcl_cot, cosmo_cot_force = force_adj(a_nbody[-1], ptcl, ptcl_cot, cosmo, conf)
    cosmo_cot = {'Omega_m': np.float64(0), 'growth': np.zeros_like(cosmo.growth)}
    for a_prev, a_next in zip(a_nbody[:0:-1], a_nbody[-2::-1]):
        ptcl, ptcl_cot, cosmo_cot, cosmo_cot_force = integrate_adj(
            a_prev, a_next, ptcl, ptcl_cot, cosmo, cosmo_cot, cosmo_cot_force, conf)
    return ptcl, ptcl_cot, cosmo_cot
