"""Oracle FFT Poisson solver: restates ``pmwd/pm_util.py:159-199,236-344`` and
``pmwd/gravity.py:9-72``.  TEST INFRASTRUCTURE ONLY.

FFTs go through ``scipy.fft`` (pocketfft, native float32/complex64 -- same family
as the DUCC FFT that JAX uses on CPU).
"""
import numpy as np
import scipy.fft as sfft

from . import pm

_WORKERS = -1


def fftfreq(shape, spacing, dtype=np.float64, sparse=True):
    """Angular wavevectors (``pmwd/pm_util.py:159-199``): computed in float64, then
    cast to ``dtype``; last axis uses ``rfftfreq``."""
    period = 1.0
    if spacing is not None:
        period = 2 * np.pi / spacing
    kvec = []
    for s in shape[:-1]:
        kvec.append((np.fft.fftfreq(s) * period).astype(dtype))
    kvec.append((np.fft.rfftfreq(shape[-1]) * period).astype(dtype))
    return np.meshgrid(*kvec, sparse=sparse, indexing='ij')


def fftfwd(f, shape=None, axes=None, norm=None):
    """``pmwd/pm_util.py:236-289``."""
    f = np.asarray(f)
    if not np.isrealobj(f):
        raise ValueError('input field must be real')
    if norm in {None, 'backward', 'ortho', 'forward'}:
        return sfft.rfftn(f, s=shape, axes=axes, norm=norm, workers=_WORKERS)
    d = f.ndim
    if shape is not None:
        d = len(shape)
    if axes is not None:
        d = len(axes)
    out = sfft.rfftn(f, s=shape, axes=axes, norm='backward', workers=_WORKERS)
    return (norm ** d * out).astype(out.dtype)


def fftinv(f, shape=None, axes=None, norm=None):
    """``pmwd/pm_util.py:292-344``."""
    f = np.asarray(f)
    if not np.iscomplexobj(f):
        raise ValueError('input field must be Hermitian complex')
    if axes is None and shape is not None:
        axes = tuple(range(-len(shape), 0))
    if norm in {None, 'backward', 'ortho', 'forward'}:
        return sfft.irfftn(f, s=shape, axes=axes, norm=norm, workers=_WORKERS)
    d = f.ndim
    if shape is not None:
        d = len(shape)
    if axes is not None:
        d = len(axes)
    out = sfft.irfftn(f, s=shape, axes=axes, norm='backward', workers=_WORKERS)
    return (norm ** -d * out).astype(out.dtype)


def laplace(kvec, src, cosmo=None):
    """``pmwd/gravity.py:9-16``: ``pot = where(k2 != 0, -src / k2, 0)``."""
    k2 = sum(k ** 2 for k in kvec)
    with np.errstate(divide='ignore', invalid='ignore'):
        pot = np.where(k2 != 0, -src / k2, 0)
    return pot.astype(src.dtype)


def neg_grad(k, pot, spacing):
    """``pmwd/gravity.py:37-44``: ``-ik * pot`` with the Nyquist planes zeroed."""
    nyquist = np.pi / spacing
    eps = nyquist * np.finfo(k.dtype).eps
    neg_ik = np.where(np.abs(np.abs(k) - nyquist) <= eps, 0, -1j * k)
    return (neg_ik * pot).astype(pot.dtype)


def rho_to_force(dens, conf, Omega_m):
    """Mesh part of ``gravity`` (``pmwd/gravity.py:49-65``): density mesh -> the
    three force meshes ``F_i`` (before the gather)."""
    fdt = conf.float_dtype
    kvec = fftfreq(conf.mesh_shape, conf.cell_size, dtype=fdt)
    dens = dens - fdt.type(1)                                  # gravity.py:52
    dens = dens * fdt.type(1.5 * np.float64(Omega_m))          # gravity.py:54
    dens = fftfwd(dens.astype(fdt))                            # gravity.py:56
    pot = laplace(kvec, dens)                                  # gravity.py:58
    out = []
    for k in kvec:
        grad = neg_grad(k, pot, conf.cell_size)                # gravity.py:62
        grad = fftinv(grad, shape=conf.mesh_shape)             # gravity.py:64
        out.append(grad.astype(fdt))
    return out


def gravity(pmid, disp, Omega_m, conf):
    """``pmwd/gravity.py:47-72``: accelerations ``(ptcl_num, dim)``."""
    dens = pm.scatter(pmid, disp, conf)                        # gravity.py:51
    forces = rho_to_force(dens, conf, Omega_m)
    acc = [pm.gather(pmid, disp, conf, F) for F in forces]     # gravity.py:67
    return np.stack(acc, axis=-1)                              # gravity.py:70


def gravity_vjp(pmid, disp, Omega_m, conf, acc_cot):
    """What ``jax.vjp(gravity)`` evaluates at ``pmwd/nbody.py:111-116``: the chain
    of the reference's own VJP rules in reverse order --

    ``_gather_bwd`` x3 (``gather.py:123-142``), transpose of ``irfftn``, transpose
    of the ``neg_grad`` multiply, ``laplace_bwd`` (``gravity.py:23-32``),
    transpose of ``rfftn``, the ``1.5*Omega_m`` scaling, ``_scatter_bwd``
    (``scatter.py:128-148``).

    Returns ``(acc, disp_cot, Omega_m_cot)``.  The FFT transposes are written via
    the identity A_i^T = -A_i for the real linear map
    A_i = irfftn . (-i k_i, Nyquist zeroed) . (-1/k^2) . rfftn
    (checked against finite differences in ``tests/test_oracle_vjp.py``).
    """
    fdt = conf.float_dtype
    acc_cot = np.asarray(acc_cot, dtype=fdt)
    dens = pm.scatter(pmid, disp, conf)
    forces = rho_to_force(dens, conf, Omega_m)
    acc = np.stack([pm.gather(pmid, disp, conf, F) for F in forces], axis=-1)

    disp_cot = np.zeros_like(acc_cot)
    kvec = fftfreq(conf.mesh_shape, conf.cell_size, dtype=fdt)
    rhop_cot_k = 0
    for i, F in enumerate(forces):
        dc, mesh_cot = pm.gather_adj(pmid, disp, conf, F, acc_cot[:, i])
        disp_cot = disp_cot + dc
        # A_i^T V_i = -A_i V_i ; accumulate in k space before one inverse FFT
        Vk = fftfwd(mesh_cot)
        rhop_cot_k = rhop_cot_k - neg_grad(kvec[i], laplace(kvec, Vk), conf.cell_size)
    rhop_cot = fftinv(rhop_cot_k, shape=conf.mesh_shape).astype(fdt)
    scale = fdt.type(1.5 * np.float64(Omega_m))
    # d/dOmega_m of  dens' = 1.5 Omega_m (dens - 1)
    Om_cot = 1.5 * np.sum(rhop_cot.astype(np.float64) * (dens.astype(np.float64) - 1))
    dens_cot = rhop_cot * scale
    dc, _ = pm.scatter_adj(pmid, disp, conf, dens_cot)
    disp_cot = disp_cot + dc
    return acc, disp_cot.astype(fdt), Om_cot
