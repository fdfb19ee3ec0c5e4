"""Lagrangian perturbation theory initial conditions (``pmwd/lpt.py:13-212``).

The force-call part of LPT (``laplace``, ``neg_grad`` + inverse FFT loop, ``lpt.py:164-173,
190-208``) uses the same fused k-space kernels as the N-body force, on the particle grid;
the strain spectra ``-k_i k_j pot`` (``lpt.py:22-32``) come from ``pmwd_strain``; the 2LPT source
product (``lpt.py:40-76``) and the displacement / velocity accumulation over orders and axes
(``lpt.py:203-208``) are one fused pass each (``csrc/lpt.cu``) with hand-written VJP kernels.
FFTs go through cuFFT (``torch.fft``); the function is differentiable (the reference uses JAX AD
with rematerialisation, ``lpt.py:136-137``): every non-FFT node of the graph is one of this
library's kernels with its own backward kernel.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .boltzmann import growth
from .cosmology import E2
from .gravity import laplace, neg_grad
from .particles import Particles
from .pm_util import fftfreq, fftfwd, fftinv


class _Strain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pot, shape, spacing, i, j):
        pot = pot.contiguous()
        out = torch.empty_like(pot)
        with torch.cuda.device(pot.device):
            _lib.check(_lib.lib().pmwd_strain(
                _lib.stream_ptr(pot.device), len(shape), _lib.shape_arr(shape), float(spacing),
                i, j, _lib.ptr(pot), _lib.ptr(out)), 'pmwd_strain')
        ctx.meta = (shape, spacing, i, j)
        return out

    @staticmethod
    def backward(ctx, cot):
        shape, spacing, i, j = ctx.meta          # real multiplier: self-adjoint
        return _Strain.apply(cot, shape, spacing, i, j), None, None, None, None


def _strain(kvec, i, j, pot, conf):
    """LPT strain component (``pmwd/lpt.py:13-37``); Nyquist planes zeroed when i != j."""
    if pot.is_cuda and pot.dtype == torch.complex64 and getattr(kvec, 'shape', None) is not None:
        strain = _Strain.apply(pot, kvec.shape, conf.ptcl_spacing, i, j)
    else:
        _lib.require_cuda(pot)     # untagged kvec: elementwise on the GPU; there is no CPU path
        k_i, k_j = kvec[i], kvec[j]
        nyquist = torch.pi / conf.ptcl_spacing
        eps = nyquist * torch.finfo(conf.float_dtype).eps
        if i != j:
            k_i = torch.where((k_i.abs() - nyquist).abs() <= eps, torch.zeros_like(k_i), k_i)
            k_j = torch.where((k_j.abs() - nyquist).abs() <= eps, torch.zeros_like(k_j), k_j)
        strain = -k_i * k_j * pot
    strain = fftinv(strain, shape=conf.ptcl_grid_shape)
    return strain.to(conf.float_dtype)


def _ptr_arr(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _Source2(torch.autograd.Function):
    """``L = s00 s22 + s00 s11 + s11 s22 - s01^2 - s02^2 - s12^2`` (``lpt.py:47-74``, m == n) in one pass."""

    @staticmethod
    def forward(ctx, *s):
        s = [t.contiguous() for t in s]
        L = torch.empty_like(s[0])
        with torch.cuda.device(L.device):
            _lib.check(_lib.lib().pmwd_lpt_source2(_lib.stream_ptr(L.device), L.numel(), _ptr_arr(s), _lib.ptr(L)),
                       'pmwd_lpt_source2')
        ctx.save_for_backward(*s)
        return L

    @staticmethod
    def backward(ctx, Lc):
        s = ctx.saved_tensors
        Lc = Lc.contiguous()
        out = [torch.empty_like(t) for t in s]
        with torch.cuda.device(Lc.device):
            _lib.check(_lib.lib().pmwd_lpt_source2_vjp(_lib.stream_ptr(Lc.device), Lc.numel(), _ptr_arr(s),
                                                       _lib.ptr(Lc), _ptr_arr(out)), 'pmwd_lpt_source2_vjp')
        return tuple(out)


class _Displace(torch.autograd.Function):
    """``disp_a = (disp0_a + D1 g1_a) + D2 g2_a`` and the same for ``vel`` with ``a^2 H D'`` (``lpt.py:203-208``),
    all orders and axes in one pass; ``D*, V*`` are 0-dim tensors (float64, any device) whose cotangents are
    the float64 sums of the VJP kernel."""

    @staticmethod
    def forward(ctx, disp0, vel0, D1, V1, D2, V2, *g):
        n = disp0.shape[0]
        g = [t.contiguous() for t in g]
        second = len(g) == 6
        f = [float(np.float32(float(x))) if x is not None else 0.0 for x in (D1, V1, D2, V2)]
        disp, vel = torch.empty_like(disp0), torch.empty_like(vel0)
        disp0, vel0 = disp0.contiguous(), vel0.contiguous()
        with torch.cuda.device(disp.device):
            _lib.check(_lib.lib().pmwd_lpt_displace(
                _lib.stream_ptr(disp.device), n, _lib.ptr(disp0), _lib.ptr(vel0), _ptr_arr(g[:3]),
                _ptr_arr(g[3:]) if second else None, f[0], f[1], f[2], f[3], _lib.ptr(disp), _lib.ptr(vel)),
                'pmwd_lpt_displace')
        ctx.save_for_backward(*g)
        ctx.meta = (f, second, [(x.dtype, x.device) if isinstance(x, torch.Tensor) else None for x in (D1, V1, D2, V2)])
        return disp, vel

    @staticmethod
    def backward(ctx, dc, vc):
        g = ctx.saved_tensors
        f, second, meta = ctx.meta
        dev = g[0].device
        n = g[0].numel()
        dc = (dc if dc is not None else torch.zeros((n, 3), device=dev)).contiguous()
        vc = (vc if vc is not None else torch.zeros((n, 3), device=dev)).contiguous()
        gc = [torch.empty_like(t) for t in g]
        sums = torch.zeros(4, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pmwd_lpt_displace_vjp(
                _lib.stream_ptr(dev), n, _lib.ptr(dc), _lib.ptr(vc), _ptr_arr(g[:3]),
                _ptr_arr(g[3:]) if second else None, f[0], f[1], f[2], f[3], _ptr_arr(gc[:3]),
                _ptr_arr(gc[3:]) if second else None, _lib.ptr(sums)), 'pmwd_lpt_displace_vjp')
        scal = [sums[k].to(device=m[1], dtype=m[0]) if m is not None else None for k, m in enumerate(meta)]
        return (dc, vc, *scal, *gc)


def _L(kvec, pot_m, pot_n, conf):
    """2LPT source (``pmwd/lpt.py:40-76``).  Same terms in the same summation order as the
    reference, but every strain component is transformed once and reused (6 inverse FFTs instead
    of 9 for ``pot_n is None``, 12 instead of 15 otherwise)."""
    m_eq_n = pot_n is None
    if m_eq_n and conf.dim == 3 and pot_m.is_cuda:
        # six strain fields, one fused product pass (csrc/lpt.cu), same terms in the same order
        st = [_strain(kvec, i, j, pot_m, conf).contiguous()
              for i, j in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))]
        return _Source2.apply(*st)
    if m_eq_n:
        pot_n = pot_m
    cache = {}

    def strain(which, pot, i, j):
        key = (which if not m_eq_n else 'm', i, j)
        if key not in cache:
            cache[key] = _strain(kvec, i, j, pot, conf)
        return cache[key]

    L = torch.zeros(conf.ptcl_grid_shape, dtype=conf.float_dtype, device=pot_m.device)
    for i in range(conf.dim):
        strain_m = strain('m', pot_m, i, i)
        for j in range(conf.dim - 1, i, -1):
            L = L + strain_m * strain('n', pot_n, j, j)
        if not m_eq_n:
            for j in range(i - 1, -1, -1):
                L = L + strain_m * strain('n', pot_n, j, j)
    if not m_eq_n:
        L = L * 0.5
    for i in range(conf.dim - 1):
        for j in range(i + 1, conf.dim):
            strain_m = strain('m', pot_m, i, j)
            strain_n = strain_m
            if not m_eq_n:
                strain_n = strain('n', pot_n, j, i)
            L = L - strain_m * strain_n
    return L


def lpt(modes, cosmo, conf):
    """Lagrangian perturbation theory at ``conf.lpt_order`` (``pmwd/lpt.py:136-212``).

    Returns ``(ptcl, obsvbl)`` with ``obsvbl = None``.
    """
    if conf.dim not in (1, 2, 3):
        raise ValueError(f'dim={conf.dim} not supported')
    if conf.lpt_order not in (0, 1, 2, 3):
        raise ValueError(f'lpt_order={conf.lpt_order} not supported')

    _lib.require_cuda(modes)
    dev = modes.device
    modes = modes / conf.ptcl_cell_vol
    kvec = fftfreq(conf.ptcl_grid_shape, conf.ptcl_spacing, dtype=conf.float_dtype, device=dev)

    pot = []
    if conf.lpt_order > 0:
        pot_1 = laplace(kvec, modes, cosmo)
        pot.append(pot_1)
    if conf.lpt_order > 1:
        src_2 = _L(kvec, pot_1, None, conf)
        src_2 = fftfwd(src_2)
        pot.append(laplace(kvec, src_2, cosmo))
    if conf.lpt_order > 2:
        raise NotImplementedError('TODO')

    a = conf.a_start
    ptcl = Particles.gen_grid(conf, vel=True, device=dev)
    disp = [ptcl.disp[:, i] for i in range(conf.dim)]
    vel = [ptcl.vel[:, i] for i in range(conf.dim)]

    if conf.dim == 3 and conf.lpt_order in (1, 2):
        # all orders and axes accumulated into the (N, 3) arrays in one fused pass
        fac, grads = [], []
        for order in range(1, 1 + conf.lpt_order):
            D = growth(a, cosmo, conf, order=order)
            dD_dlna = growth(a, cosmo, conf, order=order, deriv=1)
            fac += [D, a ** 2 * torch.sqrt(E2(a, cosmo)) * dD_dlna]
            for k in kvec:
                grad = neg_grad(k, pot[order - 1], conf.ptcl_spacing)
                grads.append(fftinv(grad, shape=conf.ptcl_grid_shape).to(conf.float_dtype).reshape(-1))
        fac += [None] * (4 - len(fac))
        fac = [torch.as_tensor(x) if x is not None else None for x in fac]
        d, v = _Displace.apply(ptcl.disp, ptcl.vel, *fac, *grads)
        return ptcl.replace(disp=d, vel=v), None

    for order in range(1, 1 + conf.lpt_order):
        D = growth(a, cosmo, conf, order=order)
        dD_dlna = growth(a, cosmo, conf, order=order, deriv=1)
        a2HDp = a ** 2 * torch.sqrt(E2(a, cosmo)) * dD_dlna
        D = D.to(conf.float_dtype).to(dev)
        a2HDp = a2HDp.to(conf.float_dtype).to(dev)
        for i, k in enumerate(kvec):
            grad = neg_grad(k, pot[order - 1], conf.ptcl_spacing)
            grad = fftinv(grad, shape=conf.ptcl_grid_shape).to(conf.float_dtype)
            grad = grad.reshape(-1)
            disp[i] = disp[i] + D * grad
            vel[i] = vel[i] + a2HDp * grad

    ptcl = ptcl.replace(disp=torch.stack(disp, dim=-1), vel=torch.stack(vel, dim=-1))
    return ptcl, None
