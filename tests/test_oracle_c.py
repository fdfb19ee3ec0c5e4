"""The compiled CPU baseline (oracle/cpm.c, C + OpenMP) against the NumPy oracle it restates:
bit-identical where the operation order is the same, summation-order tolerance for the threaded
scatter.  No GPU."""
import shutil

import numpy as np
import pytest

import oracle as O
from oracle import cpm, pm

pytestmark = pytest.mark.skipif(shutil.which('gcc') is None, reason='gcc not available')


def _state(n=12, sigma=1.7, seed=0, mesh=2):
    conf = O.Conf(1., (n, n + 2, n - 2), mesh_shape=mesh)
    rng = np.random.default_rng(seed)
    pmid = np.ascontiguousarray(O.gen_grid(conf)[0])
    disp = (sigma * rng.standard_normal((conf.ptcl_num, 3))).astype(np.float32)
    return conf, pmid, disp


def test_builds_and_reports_threads():
    cpm.build(force=True)
    assert cpm.max_threads() >= 1


@pytest.mark.parametrize('mesh', [2, 1])
def test_scatter_single_thread_is_bit_exact(mesh):
    conf, pmid, disp = _state(mesh=mesh)
    want = pm.scatter(pmid, disp, conf)
    got = cpm.scatter(pmid, disp, conf, threads=1)
    np.testing.assert_array_equal(got, want)


def test_scatter_threaded_differs_by_summation_order_only():
    conf, pmid, disp = _state(n=16, sigma=4.0)
    want = pm.scatter(pmid, disp, conf).astype(np.float64)
    got = cpm.scatter(pmid, disp, conf, threads=4).astype(np.float64)
    assert abs(got.sum() - want.sum()) <= 1e-6 * want.sum()
    np.testing.assert_allclose(got, want, rtol=0, atol=4e-6 * want.max())


def test_gather_is_bit_exact():
    conf, pmid, disp = _state(sigma=3.0)
    rng = np.random.default_rng(1)
    meshes = [rng.standard_normal(conf.mesh_shape).astype(np.float32) for _ in range(3)]
    want = np.stack([pm.gather(pmid, disp, conf, m) for m in meshes], axis=-1)
    np.testing.assert_array_equal(cpm.gather(pmid, disp, conf, meshes, threads=3), want)
    np.testing.assert_array_equal(cpm.gather(pmid, disp, conf, meshes[:1], threads=1), want[:, 0])


def test_large_displacements_wrap_like_int16():
    conf, pmid, disp = _state(n=8, sigma=40.0)       # many box lengths
    np.testing.assert_array_equal(cpm.scatter(pmid, disp, conf, threads=1), pm.scatter(pmid, disp, conf))


@pytest.mark.parametrize('shape', [(8, 6, 10), (6, 8, 9)])
def test_kspace_force_is_bit_exact(shape):
    """laplace + neg_grad (even and odd last axis: Nyquist planes exist / do not exist)."""
    conf = O.Conf(0.7, shape, mesh_shape=1)
    rng = np.random.default_rng(2)
    dens = (1 + 0.3 * rng.standard_normal(shape)).astype(np.float32)
    want = O.rho_to_force(dens.copy(), conf, 0.3)
    got = cpm.rho_to_force(dens.copy(), conf, 0.3, threads=2)
    for a in range(3):
        np.testing.assert_array_equal(got[a], want[a])


def test_step_matches_numpy_oracle():
    conf = O.Conf(1., (16, 16, 16), mesh_shape=2, a_nbody_maxstep=0.25)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    ptcl = O.lpt(O.linear_modes(O.white_noise(0, conf), cosmo, conf), cosmo, conf)
    a = conf.a_nbody
    ref = O.nbody_step(a[0], a[1], O.nbody_init(a[0], ptcl, cosmo, conf), cosmo, conf)
    one = cpm.nbody_step(a[0], a[1], cpm.nbody_init(a[0], ptcl, cosmo, conf, threads=1), cosmo, conf, threads=1)
    for k in ('disp', 'vel', 'acc'):
        np.testing.assert_array_equal(one[k], ref[k])       # same operation order -> same bits
    many = cpm.nbody_step(a[0], a[1], cpm.nbody_init(a[0], ptcl, cosmo, conf), cosmo, conf)
    for k in ('disp', 'vel', 'acc'):
        scale = np.abs(ref[k]).max()
        np.testing.assert_allclose(many[k], ref[k], rtol=0, atol=2e-5 * scale)


def test_gravity_vjp_is_bit_exact_on_one_thread():
    """The compiled adjoint force (scatter of pi, weight-gradient gathers, transposed k-space
    algebra) against oracle.gravity_vjp: same float32 operation order -> same bits."""
    conf, pmid, disp = _state(n=10, sigma=2.5)
    rng = np.random.default_rng(3)
    pi = rng.standard_normal(disp.shape).astype(np.float32)
    acc, dcot, om = O.gravity_vjp(pmid, disp, 0.3, conf, pi)
    acc1, dcot1, om1 = cpm.gravity_vjp(pmid, disp, 0.3, conf, pi, threads=1)
    np.testing.assert_array_equal(acc1, acc)
    np.testing.assert_array_equal(dcot1, dcot)
    np.testing.assert_allclose(om1, om, rtol=1e-12)
    accn, dcotn, omn = cpm.gravity_vjp(pmid, disp, 0.3, conf, pi, threads=4)
    np.testing.assert_allclose(dcotn, dcot, rtol=0, atol=2e-5 * np.abs(dcot).max())
    np.testing.assert_allclose(omn, om, rtol=1e-5)


def test_adjoint_matches_numpy_oracle():
    """cpm.nbody_adj (compiled reverse-time adjoint) vs oracle.nbody_adj on a 4-step 16^3 run."""
    conf = O.Conf(1., (16, 16, 16), mesh_shape=2, a_nbody_maxstep=0.25)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    ptcl = O.lpt(O.linear_modes(O.white_noise(0, conf), cosmo, conf), cosmo, conf)
    final = O.nbody(ptcl, cosmo, conf)
    rng = np.random.default_rng(4)
    cot = dict(disp=rng.standard_normal(final['disp'].shape).astype(np.float32),
               vel=rng.standard_normal(final['disp'].shape).astype(np.float32),
               acc=np.zeros_like(final['disp']))
    p0, c0, cc0 = O.nbody_adj(dict(final), dict(cot), cosmo, conf)
    p1, c1, cc1 = cpm.nbody_adj(dict(final), dict(cot), cosmo, conf, threads=1)
    for k in ('disp', 'vel', 'acc'):
        np.testing.assert_array_equal(p1[k], p0[k])
        np.testing.assert_array_equal(c1[k], c0[k])
    # the dot products are float64 here, float32 in NumPy: last digits of the cosmology cotangents
    np.testing.assert_allclose(cc1['Omega_m'], cc0['Omega_m'], rtol=1e-5)
    np.testing.assert_allclose(cc1['growth'], cc0['growth'], rtol=1e-4, atol=1e-6 * np.abs(cc0['growth']).max())
