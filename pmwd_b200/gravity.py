"""FFT Poisson solver with the reference's API (``pmwd/gravity.py:9-72``).

``gravity`` in 3-D runs the fused pipeline ``pmwd_force`` (scatter -> cuFFT R2C -> one
k-space kernel -> 3x cuFFT C2R -> one 3-mesh gather); its VJP is ``pmwd_force_adj``.
``laplace`` / ``neg_grad`` are the standalone k-space kernels used by LPT.
"""
import ctypes as C

import torch

from . import _lib
from .pm_util import fftfreq, fftfwd, fftinv
from .scatter import scatter, make_desc, _prep_ptcl
from .gather import gather

_workspaces = {}


def _workspace(dev, nbytes):
    """One persistent device workspace per (device, CUDA stream), grown on demand: memory is laid out once
    and reused by every force evaluation enqueued on that stream (180 GB of HBM3e make this cheap); force
    evaluations on different streams of one device get different workspaces and cannot race on the mesh
    buffers.  A buffer is allocated while its stream is current, so the caching allocator's stream-ordered
    reuse makes dropping the old one on growth safe."""
    dev = torch.device(dev)
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        _workspaces.pop(key, None)
        buf = None
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
    return buf


def release_workspaces():
    _workspaces.clear()


def _real_shape_of(kvec, spec):
    """Real-space grid shape behind a half-spectrum, from the tagged wavevectors."""
    shape = getattr(kvec, 'shape', None)
    if shape is None:
        return None
    expect = tuple(shape[:-1]) + (shape[-1] // 2 + 1,)
    return shape if tuple(spec.shape) == expect else None


class _Laplace(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, shape, spacing):
        src = src.contiguous()
        pot = torch.empty_like(src)
        with torch.cuda.device(src.device):
            _lib.check(_lib.lib().pmwd_laplace(
                _lib.stream_ptr(src.device), len(shape), _lib.shape_arr(shape), float(spacing),
                _lib.ptr(src), _lib.ptr(pot)), 'pmwd_laplace')
        ctx.meta = (shape, spacing)
        return pot

    @staticmethod
    def backward(ctx, pot_cot):
        # laplace_bwd (gravity.py:23-32): the operator is real and diagonal -> self-adjoint
        shape, spacing = ctx.meta
        return _Laplace.apply(pot_cot, shape, spacing), None, None


def laplace(kvec, src, cosmo=None):
    """Laplace kernel in Fourier space (``pmwd/gravity.py:9-34``)."""
    shape = _real_shape_of(kvec, src)
    if shape is not None and src.is_cuda and src.dtype == torch.complex64:
        return _Laplace.apply(src, shape, kvec.spacing if kvec.spacing is not None else 2 * torch.pi)
    # untagged wavevectors (not produced by fftfreq): plain elementwise evaluation on the GPU
    _lib.require_cuda(src)
    k2 = sum(k ** 2 for k in kvec)
    safe = torch.where(k2 != 0, k2, torch.ones_like(k2))
    return torch.where(k2 != 0, -src / safe, torch.zeros_like(src))


class _NegGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pot, shape, spacing, axis, sign):
        pot = pot.contiguous()
        out = torch.empty_like(pot)
        with torch.cuda.device(pot.device):
            _lib.check(_lib.lib().pmwd_neg_grad(
                _lib.stream_ptr(pot.device), len(shape), _lib.shape_arr(shape), float(spacing),
                axis, _lib.ptr(pot), _lib.ptr(out)), 'pmwd_neg_grad')
        ctx.meta = (shape, spacing, axis, sign)
        return out if sign > 0 else -out

    @staticmethod
    def backward(ctx, cot):
        # out = m(k) pot with m = -ik (masked): the VJP multiplies by conj(m) = -m
        shape, spacing, axis, sign = ctx.meta
        return _NegGrad.apply(cot, shape, spacing, axis, -sign), None, None, None, None


def neg_grad(k, pot, spacing):
    """``-ik * pot`` with the Nyquist planes zeroed (``pmwd/gravity.py:37-44``)."""
    meta = getattr(k, '_pmwd_meta', None)
    if meta is not None and pot.is_cuda and pot.dtype == torch.complex64:
        shape, kspacing, axis = meta
        expect = tuple(shape[:-1]) + (shape[-1] // 2 + 1,)
        if tuple(pot.shape) == expect and kspacing == spacing:
            return _NegGrad.apply(pot, shape, spacing, axis, 1)
    _lib.require_cuda(pot)     # untagged k: elementwise on the GPU; there is no CPU path
    nyquist = torch.pi / spacing
    eps = nyquist * torch.finfo(k.dtype).eps
    neg_ik = torch.where((k.abs() - nyquist).abs() <= eps, torch.zeros_like(k), k) * (-1j)
    return neg_ik * pot


class _Gravity(torch.autograd.Function):
    """3-D fast path: ``gravity`` (``pmwd/gravity.py:47-72``) and its VJP as evaluated by
    ``force_adj`` (``pmwd/nbody.py:108-118``)."""

    @staticmethod
    def forward(ctx, disp, Omega_m, pmid, conf):
        pmid, disp = _prep_ptcl(pmid, disp, conf)
        acc = torch.empty_like(disp)
        force_into(pmid, disp, float(Omega_m), conf, acc)
        ctx.save_for_backward(pmid, disp, acc)
        ctx.meta = (conf, float(Omega_m), Omega_m)
        return acc

    @staticmethod
    def backward(ctx, acc_cot):
        pmid, disp, acc = ctx.saved_tensors
        conf, Om, Om_in = ctx.meta
        pi = acc_cot.to(conf.float_dtype).contiguous()
        alpha = torch.empty_like(disp)
        acc2 = torch.empty_like(disp)
        force_adj_into(pmid, disp, Om, conf, pi, acc2, alpha)
        Om_cot = None
        if isinstance(Om_in, torch.Tensor) and ctx.needs_input_grad[1]:
            # gravity is linear in Omega_m (gravity.py:54): zeta = sum(pi . acc) / Omega_m
            Om_cot = ((pi.double() * acc.double()).sum() / Om).to(Om_in.dtype).to(Om_in.device)
        return alpha, Om_cot, None, None


def _force_desc(pmid, conf):
    return make_desc(conf, pmid, conf.mesh_shape, 1, 0, None)


def _mode(conf):
    return _lib.SCATTER_DETERMINISTIC if conf.scatter_mode == 'deterministic' else _lib.SCATTER_ATOMIC


def force_into(pmid, disp, Omega_m, conf, acc, kick_vel=None, kick_factor=0.0, sweep=None):
    """Enqueue ``pmwd_force``: ``acc <- gravity`` (and ``kick_vel += acc * kick_factor``)."""
    dev = disp.device
    desc = _force_desc(pmid, conf)
    mode = _mode(conf)
    lib = _lib.lib()
    nbytes = lib.pmwd_force_workspace_bytes(C.byref(desc), 0, mode)
    ws = _workspace(dev, nbytes)
    ctx = _lib.Context.get(dev).reserve(conf.mesh_shape)
    with torch.cuda.device(dev):
        _lib.check(lib.pmwd_force(
            ctx.handle, _lib.stream_ptr(dev), C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp),
            float(Omega_m), _lib.ptr(acc), _lib.ptr(kick_vel), float(kick_factor), mode,
            _lib.ptr(ws), ws.numel(), sweep), 'pmwd_force')


def force_kdk_into(pmid, disp, Omega_m, conf, acc, vel, K2, K1_next, D_next, sweep=None):
    """Enqueue ``pmwd_force_kdk``: force at ``disp``, trailing half-kick ``K2``, then the next
    step's leading half-kick ``K1_next`` and drift ``D_next`` -- one call per KDK step."""
    dev = disp.device
    desc = _force_desc(pmid, conf)
    mode = _mode(conf)
    lib = _lib.lib()
    nbytes = lib.pmwd_force_workspace_bytes(C.byref(desc), 0, mode)
    ws = _workspace(dev, nbytes)
    ctx = _lib.Context.get(dev).reserve(conf.mesh_shape)
    with torch.cuda.device(dev):
        _lib.check(lib.pmwd_force_kdk(
            ctx.handle, _lib.stream_ptr(dev), C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp),
            float(Omega_m), _lib.ptr(acc), _lib.ptr(vel), float(K2), float(K1_next), float(D_next), mode,
            _lib.ptr(ws), ws.numel(), sweep), 'pmwd_force_kdk')


def force_adj_into(pmid, disp, Omega_m, conf, pi, acc, alpha, sweep=None):
    """Enqueue ``pmwd_force_adj``: ``acc <- gravity``, ``alpha <- VJP_disp(gravity)(pi)``."""
    dev = disp.device
    desc = _force_desc(pmid, conf)
    mode = _mode(conf)
    lib = _lib.lib()
    nbytes = lib.pmwd_force_workspace_bytes(C.byref(desc), 1, mode)
    ws = _workspace(dev, nbytes)
    ctx = _lib.Context.get(dev).reserve(conf.mesh_shape)
    with torch.cuda.device(dev):
        _lib.check(lib.pmwd_force_adj(
            ctx.handle, _lib.stream_ptr(dev), C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp),
            float(Omega_m), _lib.ptr(pi), _lib.ptr(acc), _lib.ptr(alpha), mode,
            _lib.ptr(ws), ws.numel(), sweep), 'pmwd_force_adj')


def _gravity_general(ptcl, cosmo, conf):
    """Any-dim composition exactly as written in ``pmwd/gravity.py:47-72`` out of the
    differentiable building blocks (1-D / 2-D meshes, non-int16 pmid)."""
    kvec = fftfreq(conf.mesh_shape, conf.cell_size, dtype=conf.float_dtype, device=ptcl.disp.device)
    dens = scatter(ptcl, conf)
    dens = dens - 1
    Om = cosmo.Omega_m
    Om = Om.to(device=dens.device, dtype=conf.float_dtype) if isinstance(Om, torch.Tensor) else Om
    dens = dens * (1.5 * Om)
    dens = fftfwd(dens)
    pot = laplace(kvec, dens, cosmo)
    acc = []
    for k in kvec:
        grad = neg_grad(k, pot, conf.cell_size)
        grad = fftinv(grad, shape=conf.mesh_shape).to(conf.float_dtype)
        acc.append(gather(ptcl, conf, grad))
    return torch.stack(acc, dim=-1)


def gravity(a, ptcl, cosmo, conf):
    """Gravitational accelerations of particles in [H_0^2] (``pmwd/gravity.py:47-72``)."""
    _lib.require_cuda(ptcl.pmid, ptcl.disp)
    if conf.dim == 3 and ptcl.pmid.dtype == torch.int16 and ptcl.disp.is_cuda:
        return _Gravity.apply(ptcl.disp, cosmo.Omega_m, ptcl.pmid, conf)
    return _gravity_general(ptcl, cosmo, conf)
