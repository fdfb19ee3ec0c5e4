"""ctypes front end of ``oracle/cpm.c``: the compiled (C + OpenMP) restatement of one
forward KDK step, for the CPU baseline.  TEST INFRASTRUCTURE ONLY.

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

from .nbody import drift_factor, kick_factor
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
        for f in (L.cpm_scatter, L.cpm_gather, L.cpm_kspace_force, L.cpm_contrast, L.cpm_axpy):
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
