"""CPU checks of the power-spectrum estimator's host logic and per-mode arithmetic.

``pmwd_b200/csrc/powspec.cuh`` holds the per-mode function the CUDA kernel uses; here it is
compiled for the host (``tests/host/powspec_emul.cc``), applied to every mode of a NumPy
spectrum, and the resulting P(k) is held to the oracle's ``powspec`` (``pmwd/spec_util.py:50-147``).
"""
import ctypes as C
import math
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
from oracle.spec import _getbins as oracle_getbins

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('shape', [(8, 8, 8), (12, 10, 9), (16, 16, 16)])
@pytest.mark.parametrize('bins', [1j / 3, 1j / 2, 1, 2.5, (0, 0.1, 0.2, 0.35, 0.6, 0.9)])
@pytest.mark.parametrize('cut_nyq', [True, False])
def test_getbins_matches_oracle(shape, bins, cut_nyq):
    from pmwd_b200.spec_util import _getbins
    bnum, bcut, edges, right = _getbins(shape, bins, cut_nyq)
    onum, ocut, oedges, oright = oracle_getbins(shape, bins, cut_nyq)
    assert (bnum, bcut, right) == (onum, ocut, oright)
    np.testing.assert_allclose(edges, oedges, rtol=1e-15)


@pytest.fixture(scope='module')
def emul(tmp_path_factory):
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    so = tmp_path_factory.mktemp('ps') / 'libpowspec_emul.so'
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off',
                    '-I' + os.path.join(ROOT, 'pmwd_b200', 'csrc'),
                    os.path.join(ROOT, 'tests', 'host', 'powspec_emul.cc'), '-o', str(so)], check=True)
    return C.CDLL(str(so))


def _emulated_powspec(emul, f, spacing, bins, g, deconv, cut_zero, cut_nyq):
    from pmwd_b200.spec_util import _getbins
    shape = f.shape
    bnum, bcut, edges, right = _getbins(shape, bins, cut_nyq)
    edges = np.asarray(edges, dtype=np.float64)
    fk = np.ascontiguousarray(O.fftfwd(f)).astype(np.complex64)
    gk = None if g is None else np.ascontiguousarray(O.fftfwd(g)).astype(np.complex64)
    out = np.zeros((4, bnum + 1))
    vp = C.c_void_p
    emul.ps_emul((C.c_int * 3)(*shape), fk.ctypes.data_as(vp), None if gk is None else gk.ctypes.data_as(vp),
                 int(deconv is not None), C.c_double(deconv or 0.), edges.ctypes.data_as(vp), bnum, int(right),
                 out.ctypes.data_as(vp))
    lo = int(cut_zero)
    N = out[3, lo:bcut]
    P = out[1, lo:bcut] if g is None else out[1, lo:bcut] + 1j * out[2, lo:bcut]
    with np.errstate(invalid='ignore', divide='ignore'):
        k = out[0, lo:bcut] / N * (2 * math.pi / spacing)
        P = P / N * (spacing ** 3 / math.prod(shape))
    return k, P, N, edges[:bcut] * (2 * math.pi / spacing)


@pytest.mark.parametrize('shape', [(16, 16, 16), (12, 10, 9), (8, 14, 11)])
@pytest.mark.parametrize('bins, deconv, cross', [(1j / 3, None, False), (1, 2, False), (1j / 2, None, True),
                                                 (2.0, 1, True), ((0, 0.1, 0.25, 0.4, 0.7), None, False)])
def test_per_mode_arithmetic_vs_oracle(emul, shape, bins, deconv, cross):
    rng = np.random.default_rng(3)
    f = rng.standard_normal(shape).astype(np.float32)
    g = (0.5 * f + rng.standard_normal(shape)).astype(np.float32) if cross else None
    for cut_zero, cut_nyq in ((True, True), (False, False)):
        want = O.powspec(f, 0.7, bins=bins, g=g, deconv=deconv, cut_zero=cut_zero, cut_nyq=cut_nyq)
        got = _emulated_powspec(emul, f, 0.7, bins, g, deconv, cut_zero, cut_nyq)
        np.testing.assert_array_equal(got[2], want[2])                       # mode counts: exact
        np.testing.assert_allclose(got[3], want[3], rtol=1e-15)
        ok = want[2] > 0
        np.testing.assert_allclose(got[0][ok], want[0][ok], rtol=1e-12)
        np.testing.assert_allclose(got[1][ok], want[1][ok], rtol=2e-6, atol=1e-6 * np.abs(want[1][ok]).max())


@pytest.mark.parametrize('shape, bins, deconv', [((12, 10, 9), 1j / 3, None), ((8, 8, 8), 1, 2), ((10, 12, 8), 1j / 2, 1)])
def test_vjp_weighting_vs_finite_differences(emul, shape, bins, deconv):
    """L = sum_b c_b P_b: the field gradient assembled from the per-mode weights the kernel applies
    (twice the unnormalised inverse FFT of w_k f_k) against central differences of the float64 oracle."""
    from pmwd_b200.spec_util import _getbins
    rng = np.random.default_rng(7)
    f = rng.standard_normal(shape)
    spacing = 0.7
    bnum, bcut, edges, right = _getbins(shape, bins, True)
    edges = np.asarray(edges, dtype=np.float64)
    k, P, N, _ = O.powspec(f, spacing, bins=bins, deconv=deconv)
    c = rng.standard_normal(P.shape)
    c[N == 0] = 0

    def loss(field):
        Pb = O.powspec(field, spacing, bins=bins, deconv=deconv)[1]
        return float(np.sum(np.where(N > 0, c * Pb, 0.0)))

    # per-bin weights as pmwd_b200.spec_util hands them to the kernel: cotangent of the bin SUM
    wbin = np.zeros(bnum + 1)
    wbin[1:bcut] = np.where(N > 0, c * (spacing ** 3 / math.prod(shape)) / np.where(N > 0, N, 1), 0.0)
    fk = np.ascontiguousarray(O.fftfwd(f.astype(np.float32))).astype(np.complex64)
    out = np.empty_like(fk)
    vp = C.c_void_p
    emul.ps_weight_emul((C.c_int * 3)(*shape), fk.ctypes.data_as(vp), int(deconv is not None), C.c_double(deconv or 0.),
                        edges.ctypes.data_as(vp), bnum, int(right), wbin.ctypes.data_as(vp), out.ctypes.data_as(vp))
    grad = 2 * np.fft.irfftn(out.astype(np.complex128), s=shape, axes=(0, 1, 2), norm='forward')
    h = 1e-5
    idx = [tuple(rng.integers(0, n) for n in shape) for _ in range(12)]
    for ix in idx:
        fp, fm = f.copy(), f.copy()
        fp[ix] += h
        fm[ix] -= h
        fd = (loss(fp) - loss(fm)) / (2 * h)
        assert abs(grad[ix] - fd) <= 2e-4 * np.abs(grad).max() + 1e-9, (ix, grad[ix], fd)
