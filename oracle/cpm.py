"""ctypes front end of ``oracle/cpm.c``: the compiled (C + OpenMP) restatement of one
forward KDK step and of one reverse-time adjoint step, for the CPU baseline and the
full-size parity tests.  TEST INFRASTRUCTURE ONLY.

``oracle/cpm.c`` repeats the float32 arithmetic of the NumPy oracle (hence of the reference,
see the citations there) with plain loops; ``tests/test_oracle_c.py`` holds it to the NumPy
oracle: single-threaded scatter, gather, k-space and kick/drift are bit-identical, the
multi-threaded scatter differs only in summation order.  FFTs stay in scipy (pocketfft,
all cores).  ``bench.py`` times this port as ``cpu_baseline`` / ``--impl reference``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .nbody import drift_factor, kick_factor, factor_grads, _cot_axpy
from .gravity import fftfwd, fftinv

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'cpm.c')
LIB = os.path.join(HERE, '_build', 'libcpm.so')
_lib = None


def build(force=False):
    """gcc -O3 -fopenmp, no fast-math, no FMA contraction (NumPy rounds every product)."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ['gcc', '-O3', '-fopenmp', '-fno-fast-math', '-ffp-contract=off', '-shared', '-fPIC',
           SRC, '-o', LIB, '-lm']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('gcc failed for oracle/cpm.c:\n' + r.stdout + r.stderr)
    return LIB


def available():
    try:
        lib()
        return True
    except (OSError, RuntimeError):
        return False


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        i64, f32, i32p, fp, i16p, vp = C.c_int64, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_float), \
            C.POINTER(C.c_int16), C.c_void_p
        L.cpm_max_threads.restype = C.c_int
        L.cpm_scatter.argtypes = [i64, vp, vp, f32, f32, i32p, vp, C.c_int]
        L.cpm_gather.argtypes = [i64, vp, vp, f32, i32p, C.c_int, vp, vp, vp, vp, C.c_int]
        L.cpm_kspace_force.argtypes = [i32p, C.c_double, vp, vp, vp, vp, C.c_int]
        L.cpm_contrast.argtypes = [i64, vp, f32, C.c_int]
        L.cpm_axpy.argtypes = [i64, vp, vp, f32, C.c_int]
        L.cpm_scatter_val.argtypes = [i64, vp, vp, vp, C.c_int, f32, i32p, vp, C.c_int]
        L.cpm_grad_gather.argtypes = [i64, vp, vp, f32, i32p, vp, vp, C.c_int, f32, vp, C.c_int]
        L.cpm_kspace_force_adj.argtypes = [i32p, C.c_double, vp, vp, vp, vp, C.c_int]
        L.cpm_dot.argtypes = [i64, vp, vp, C.c_int]
        L.cpm_dot.restype = C.c_double
        for f in (L.cpm_scatter, L.cpm_gather, L.cpm_kspace_force, L.cpm_contrast, L.cpm_axpy,
                  L.cpm_scatter_val, L.cpm_grad_gather, L.cpm_kspace_force_adj):
            f.restype = None
        _lib = L
    return _lib


def max_threads():
    return int(lib().cpm_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _shape(shape):
    return (C.c_int32 * 3)(*[int(s) for s in shape])


def _check_fast(pmid, disp, conf):
    assert pmid.dtype == np.int16 and disp.dtype == np.float32 and conf.float_dtype == np.float32
    assert pmid.shape[1] == 3 and pmid.flags.c_contiguous and disp.flags.c_contiguous


def scatter(pmid, disp, conf, threads=None):
    """``pm.scatter(pmid, disp, conf)`` (defaults: zero mesh, val = N_m / N_p)."""
    _check_fast(pmid, disp, conf)
    threads = max_threads() if threads is None else threads
    mesh = np.zeros(conf.mesh_shape, dtype=np.float32)
    val = np.float32(conf.mesh_size / conf.ptcl_num)
    lib().cpm_scatter(len(pmid), _p(pmid), _p(disp), val, np.float32(conf.cell_size), _shape(conf.mesh_shape),
                      _p(mesh), threads)
    return mesh


def gather(pmid, disp, conf, meshes, threads=None):
    """``pm.gather`` of 1..3 scalar meshes in one pass: ``(N,)`` for one mesh, ``(N, 3)`` for three."""
    _check_fast(pmid, disp, conf)
    threads = max_threads() if threads is None else threads
    meshes = [np.ascontiguousarray(m, dtype=np.float32) for m in meshes]
    nm = len(meshes)
    assert nm in (1, 3)
    out = np.empty((len(pmid), nm) if nm == 3 else (len(pmid),), dtype=np.float32)
    ptrs = [_p(m) for m in meshes] + [None] * (3 - nm)
    lib().cpm_gather(len(pmid), _p(pmid), _p(disp), np.float32(conf.cell_size), _shape(conf.mesh_shape), nm,
                     ptrs[0], ptrs[1], ptrs[2], _p(out), threads)
    return out


def rho_to_force(dens, conf, Omega_m, threads=None):
    """``gravity.rho_to_force``; ``dens`` is consumed (modified in place)."""
    threads = max_threads() if threads is None else threads
    L = lib()
    L.cpm_contrast(dens.size, _p(dens), np.float32(1.5 * np.float64(Omega_m)), threads)
    spec = np.ascontiguousarray(fftfwd(dens))
    assert spec.dtype == np.complex64
    g = [np.empty_like(spec) for _ in range(3)]
    L.cpm_kspace_force(_shape(conf.mesh_shape), float(conf.cell_size), _p(spec), _p(g[0]), _p(g[1]), _p(g[2]),
                       threads)
    return [fftinv(gi, shape=conf.mesh_shape).astype(np.float32, copy=False) for gi in g]


def gravity(pmid, disp, Omega_m, conf, threads=None):
    """``gravity.gravity`` (``pmwd/gravity.py:47-72``)."""
    dens = scatter(pmid, disp, conf, threads)
    forces = rho_to_force(dens, conf, Omega_m, threads)
    return gather(pmid, disp, conf, forces, threads)


def _axpy(y, x, f, threads):
    out = y.copy()
    lib().cpm_axpy(out.size, _p(out), _p(np.ascontiguousarray(x)), np.float32(f), threads)
    return out


def nbody_init(a, ptcl, cosmo, conf, threads=None):
    """``nbody.nbody_init``."""
    ptcl = dict(ptcl)
    ptcl['acc'] = gravity(ptcl['pmid'], ptcl['disp'], cosmo.Omega_m, conf, threads)
    return ptcl


def nbody_step(a_prev, a_next, ptcl, cosmo, conf, threads=None):
    """``nbody.integrate`` (``pmwd/nbody.py:121-140``) on the compiled kernels; step factors
    from the NumPy oracle (float64 host scalars)."""
    threads = max_threads() if threads is None else threads
    ptcl = dict(ptcl)
    D = K = 0
    a_disp = a_vel = a_acc = a_prev
    for d, k in conf.symp_splits:
        if d != 0:
            D += d
            a_disp_next = a_prev * (1 - D) + a_next * D
            f = np.float32(drift_factor(a_vel, a_disp, a_disp_next, cosmo, conf))
            ptcl['disp'] = _axpy(ptcl['disp'], ptcl['vel'], f, threads)
            a_disp = a_disp_next
            ptcl['acc'] = gravity(ptcl['pmid'], ptcl['disp'], cosmo.Omega_m, conf, threads)
            a_acc = a_disp
        if k != 0:
            K += k
            a_vel_next = a_prev * (1 - K) + a_next * K
            f = np.float32(kick_factor(a_acc, a_vel, a_vel_next, cosmo, conf))
            ptcl['vel'] = _axpy(ptcl['vel'], ptcl['acc'], f, threads)
            a_vel = a_vel_next
    return ptcl


# ----------------------------------------------------------------------------- adjoint
def gravity_vjp(pmid, disp, Omega_m, conf, acc_cot, threads=None):
    """``oracle.gravity.gravity_vjp`` (what ``jax.vjp(gravity)`` evaluates at
    ``pmwd/nbody.py:111-116``) on the compiled loops: returns ``(acc, disp_cot, Omega_m_cot)``.
    Same float32 operation order as the NumPy oracle (bit-identical on one thread, the two
    scatters' summation order aside on several)."""
    _check_fast(pmid, disp, conf)
    threads = max_threads() if threads is None else threads
    L = lib()
    n = len(pmid)
    cell = np.float32(conf.cell_size)
    shp = _shape(conf.mesh_shape)
    acc_cot = np.ascontiguousarray(acc_cot, dtype=np.float32)
    dens = scatter(pmid, disp, conf, threads)
    dens1 = dens.copy()
    forces = rho_to_force(dens1, conf, Omega_m, threads)          # consumes dens1
    acc = gather(pmid, disp, conf, forces, threads)
    disp_cot = np.zeros((n, 3), dtype=np.float32)
    spec = []
    for i in range(3):
        F = np.ascontiguousarray(forces[i])
        L.cpm_grad_gather(n, _p(pmid), _p(disp), cell, shp, _p(F), _p(acc_cot[:, i:]), 3, np.float32(0),
                          _p(disp_cot), threads)                  # gather.py:106-110
        V = np.zeros(conf.mesh_shape, dtype=np.float32)
        L.cpm_scatter_val(n, _p(pmid), _p(disp), _p(acc_cot[:, i:]), 3, cell, shp, _p(V), threads)   # :113
        spec.append(np.ascontiguousarray(fftfwd(V)))
        del V
    out = np.empty_like(spec[0])
    L.cpm_kspace_force_adj(shp, float(conf.cell_size), _p(spec[0]), _p(spec[1]), _p(spec[2]), _p(out), threads)
    del spec
    rhop_cot = fftinv(out, shape=conf.mesh_shape).astype(np.float32, copy=False)
    scale = np.float32(1.5 * np.float64(Omega_m))
    Om_cot = 1.5 * np.sum(rhop_cot.astype(np.float64) * (dens.astype(np.float64) - 1))
    dens_cot = np.ascontiguousarray(rhop_cot * scale)
    val = np.float32(conf.mesh_size / conf.ptcl_num)
    L.cpm_grad_gather(n, _p(pmid), _p(disp), cell, shp, _p(dens_cot), None, 0, val, _p(disp_cot), threads)
    return acc, disp_cot, Om_cot


def force_adj(a, ptcl, ptcl_cot, cosmo, conf, threads=None):
    """``oracle.nbody.force_adj`` (``pmwd/nbody.py:108-118``)."""
    acc, disp_cot, Om_cot = gravity_vjp(ptcl['pmid'], ptcl['disp'], cosmo.Omega_m, conf, ptcl_cot['vel'], threads)
    ptcl = dict(ptcl); ptcl['acc'] = acc
    ptcl_cot = dict(ptcl_cot); ptcl_cot['acc'] = disp_cot
    return ptcl, ptcl_cot, Om_cot


def _dot(x, y, threads):
    x = np.ascontiguousarray(x); y = np.ascontiguousarray(y)
    return lib().cpm_dot(x.size, _p(x), _p(y), threads)


def nbody_adj_step(a_prev, a_next, ptcl, ptcl_cot, cosmo, cosmo_cot, cosmo_cot_force, conf, threads=None):
    """``oracle.nbody.integrate_adj`` (``pmwd/nbody.py:143-162`` with ``kick_adj`` :80-99 and
    ``drift_adj`` :49-67 inlined) on the compiled kernels.  The two dot products per step are
    accumulated in float64 by threads (the NumPy oracle sums in float32)."""
    threads = max_threads() if threads is None else threads
    ptcl, ptcl_cot = dict(ptcl), dict(ptcl_cot)
    K = D = 0
    a_disp = a_vel = a_acc = a_prev
    for d, k in reversed(conf.symp_splits):
        if k != 0:
            K += k
            a_vel_next = a_prev * (1 - K) + a_next * K
            factor, grads = factor_grads(kick_factor, a_acc, a_vel, a_vel_next, cosmo, conf)
            f = np.float32(factor)
            ptcl['vel'] = _axpy(ptcl['vel'], ptcl['acc'], f, threads)
            ptcl_cot['disp'] = _axpy(ptcl_cot['disp'], ptcl_cot['acc'], -f, threads)
            s = _dot(ptcl_cot['vel'], ptcl['acc'], threads)
            cosmo_cot = _cot_axpy(cosmo_cot, s, grads)
            cosmo_cot = dict(cosmo_cot)
            cosmo_cot['Omega_m'] = cosmo_cot['Omega_m'] - cosmo_cot_force * np.float64(f)
            a_vel = a_vel_next
        if d != 0:
            D += d
            a_disp_next = a_prev * (1 - D) + a_next * D
            factor, grads = factor_grads(drift_factor, a_vel, a_disp, a_disp_next, cosmo, conf)
            f = np.float32(factor)
            ptcl['disp'] = _axpy(ptcl['disp'], ptcl['vel'], f, threads)
            ptcl_cot['vel'] = _axpy(ptcl_cot['vel'], ptcl_cot['disp'], -f, threads)
            s = _dot(ptcl_cot['disp'], ptcl['vel'], threads)
            cosmo_cot = _cot_axpy(cosmo_cot, s, grads)
            a_disp = a_disp_next
            ptcl, ptcl_cot, cosmo_cot_force = force_adj(a_disp, ptcl, ptcl_cot, cosmo, conf, threads)
            a_acc = a_disp
    return ptcl, ptcl_cot, cosmo_cot, cosmo_cot_force


def nbody_adj_init(a, ptcl, ptcl_cot, cosmo, conf, threads=None):
    """``pmwd/nbody.py:226-236``: returns ``(ptcl, ptcl_cot, cosmo_cot, cosmo_cot_force)``."""
    ptcl, ptcl_cot, cosmo_cot_force = force_adj(a, ptcl, ptcl_cot, cosmo, conf, threads)
    cosmo_cot = {'Omega_m': np.float64(0), 'growth': np.zeros_like(cosmo.growth)}
    return ptcl, ptcl_cot, cosmo_cot, cosmo_cot_force


def nbody_adj(ptcl, ptcl_cot, cosmo, conf, threads=None):
    """``oracle.nbody.nbody_adj`` (``pmwd/nbody.py:226-260``)."""
    a_nbody = conf.a_nbody
    ptcl, ptcl_cot, cosmo_cot, ccf = nbody_adj_init(a_nbody[-1], ptcl, ptcl_cot, cosmo, conf, threads)
    for a_prev, a_next in zip(a_nbody[:0:-1], a_nbody[-2::-1]):
        ptcl, ptcl_cot, cosmo_cot, ccf = nbody_adj_step(a_prev, a_next, ptcl, ptcl_cot, cosmo, cosmo_cot, ccf,
                                                        conf, threads)
    return ptcl, ptcl_cot, cosmo_cot
