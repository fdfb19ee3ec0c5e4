"""Structural pins of the oracle's VJPs in float64: analytic VJP vs central finite
differences / the dot-product identity <J v, w> = <v, J^T w> -- the oracle
analogue of ``pmwd/test_util.py:88-149`` and
``/root/reference/tests/pm_test.py:121-171`` (custom VJP == autodiff / numerical
VJP, incl. even/odd meshes for the Nyquist handling)."""
import numpy as np
import pytest

import oracle as O


def _setup(shape, disp_std=1.3, seed=0, mesh_ratio=1):
    conf = O.Conf(1.7, shape, mesh_shape=mesh_ratio, float_dtype=np.float64)
    pmid, disp, _, _ = O.gen_grid(conf)
    rng = np.random.default_rng(seed)
    disp = disp + disp_std * rng.standard_normal(disp.shape)
    return conf, pmid, disp, rng


def _fd(f, x, v, h=1e-6):
    return (f(x + h * v) - f(x - h * v)) / (2 * h)


@pytest.mark.parametrize('shape', [(4, 9), (7, 8), (4, 5, 6)])
def test_scatter_gather_vjp_fd(shape):
    conf, pmid, disp, rng = _setup(shape)
    dim = len(shape)
    chan = (2,)
    val = rng.standard_normal((len(pmid),) + chan)
    mesh_cot = rng.standard_normal(conf.mesh_shape + chan)
    v = rng.standard_normal(disp.shape)

    # scatter: d/d disp and d/d val
    f = lambda d: np.sum(O.scatter(pmid, d, conf, val=val) * mesh_cot)
    disp_cot, val_cot = O.scatter_adj(pmid, disp, conf, mesh_cot, val=val)
    np.testing.assert_allclose(np.sum(disp_cot * v), _fd(f, disp, v), rtol=1e-6)
    vv = rng.standard_normal(val.shape)
    g = lambda a: np.sum(O.scatter(pmid, disp, conf, val=a) * mesh_cot)
    np.testing.assert_allclose(np.sum(val_cot * vv), _fd(g, val, vv), rtol=1e-7)

    # gather: d/d disp and d/d mesh
    mesh = rng.standard_normal(conf.mesh_shape + chan)
    val_cot = rng.standard_normal((len(pmid),) + chan)
    z = np.zeros_like(val_cot)
    f = lambda d: np.sum(O.gather(pmid, d, conf, mesh, val=z) * val_cot)
    disp_cot, mesh_cot2 = O.gather_adj(pmid, disp, conf, mesh, val_cot)
    np.testing.assert_allclose(np.sum(disp_cot * v), _fd(f, disp, v), rtol=1e-6)
    mv = rng.standard_normal(mesh.shape)
    g = lambda m: np.sum(O.gather(pmid, disp, conf, m, val=z) * val_cot)
    np.testing.assert_allclose(np.sum(mesh_cot2 * mv), _fd(g, mesh, mv), rtol=1e-7)
    # gather is the transpose of scatter
    np.testing.assert_allclose(mesh_cot2, O.scatter(pmid, disp, conf, val=val_cot,
                               mesh=np.zeros_like(mesh)), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize('shape', [(4, 9), (7, 8), (4, 5, 6), (6, 6, 8)])
def test_gravity_vjp_fd(shape):
    """pm_test.py:154-171: VJP of gravity vs numerical, even and odd meshes."""
    conf, pmid, disp, rng = _setup(shape)
    acc_cot = rng.standard_normal(disp.shape)
    acc, disp_cot, Om_cot = O.gravity_vjp(pmid, disp, 0.3, conf, acc_cot)
    np.testing.assert_allclose(acc, O.gravity(pmid, disp, 0.3, conf), rtol=1e-13, atol=1e-14)
    v = rng.standard_normal(disp.shape)
    f = lambda d: np.sum(O.gravity(pmid, d, 0.3, conf) * acc_cot)
    np.testing.assert_allclose(np.sum(disp_cot * v), _fd(f, disp, v), rtol=2e-6)
    g = lambda om: np.sum(O.gravity(pmid, disp, om, conf) * acc_cot)
    np.testing.assert_allclose(Om_cot, _fd(g, 0.3, 1.0), rtol=1e-7)
    # zeta = sum(pi . acc) / Omega_m  (gravity is linear in Omega_m)
    np.testing.assert_allclose(Om_cot, np.sum(acc_cot * acc) / 0.3, rtol=1e-10)


def _nbody_obj(conf, cosmo, ptcl, w_disp, w_vel):
    out = O.nbody(ptcl, cosmo, conf)
    return np.sum(out['disp'] * w_disp) + np.sum(out['vel'] * w_vel)


def test_nbody_adj_fd():
    """pm_test.py:194-211 analogue: adjoint of the whole integration vs numerical
    derivatives (w.r.t. initial disp/vel, Omega_m and growth-table entries)."""
    conf = O.Conf(2., (6, 6, 6), mesh_shape=2, float_dtype=np.float64,
                  a_start=1 / 4, a_stop=1., a_nbody_maxstep=1 / 4)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    pmid, disp, vel, _ = O.gen_grid(conf, vel=True)
    rng = np.random.default_rng(1)
    disp = disp + 0.7 * rng.standard_normal(disp.shape)
    vel = 0.3 * rng.standard_normal(disp.shape)
    ptcl = dict(pmid=pmid, disp=disp, vel=vel)
    w_disp = rng.standard_normal(disp.shape)
    w_vel = rng.standard_normal(disp.shape)

    final = O.nbody(ptcl, cosmo, conf)
    cot = dict(disp=w_disp.copy(), vel=w_vel.copy(), acc=np.zeros_like(disp))
    back, pc, cc = O.nbody_adj(final, cot, cosmo, conf)
    # reverse-time reconstruction returns the initial state (nbody.py:263-265)
    np.testing.assert_allclose(back['disp'], disp, atol=1e-9)
    np.testing.assert_allclose(back['vel'], vel, atol=1e-9)

    v = rng.standard_normal(disp.shape)
    f = lambda d: _nbody_obj(conf, cosmo, dict(pmid=pmid, disp=d, vel=vel), w_disp, w_vel)
    np.testing.assert_allclose(np.sum(pc['disp'] * v), _fd(f, disp, v), rtol=2e-5)
    f = lambda u: _nbody_obj(conf, cosmo, dict(pmid=pmid, disp=disp, vel=u), w_disp, w_vel)
    np.testing.assert_allclose(np.sum(pc['vel'] * v), _fd(f, vel, v), rtol=2e-5)
    # cosmology leaves: direct Omega_m dependence with the growth table held fixed
    f = lambda om: _nbody_obj(conf, cosmo.replace(Omega_m=om), ptcl, w_disp, w_vel)
    np.testing.assert_allclose(cc['Omega_m'], _fd(f, cosmo.Omega_m, 1.0, h=1e-5), rtol=2e-5)
    # growth-table leaf
    tv = rng.standard_normal(cosmo.growth.shape) * 0.01
    f = lambda t: _nbody_obj(conf, cosmo.replace(growth=t), ptcl, w_disp, w_vel)
    np.testing.assert_allclose(np.sum(cc['growth'] * tv), _fd(f, cosmo.growth, tv, h=1e-5),
                               rtol=2e-5)


@pytest.mark.parametrize('order', [1, 2])
def test_lpt_vjp_fd(order):
    conf = O.Conf(1.5, (8, 8, 8), mesh_shape=2, float_dtype=np.float64, lpt_order=order)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    rng = np.random.default_rng(2)
    white = rng.standard_normal(conf.ptcl_grid_shape) * 30   # boost so 2LPT matters
    dc = rng.standard_normal((conf.ptcl_num, 3))
    vc = rng.standard_normal((conf.ptcl_num, 3))

    def f(w):
        p = O.lpt(O.linear_modes(w, cosmo, conf), cosmo, conf)
        return np.sum(p['disp'] * dc) + np.sum(p['vel'] * vc)
    g = O.lpt_vjp_modes(white, cosmo, conf, dc, vc)
    v = rng.standard_normal(white.shape)
    np.testing.assert_allclose(np.sum(g * v), _fd(f, white, v, h=1e-4), rtol=1e-6)


def test_reversibility():
    """pm_test.py:177-192: forward then reverse integration returns the start."""
    conf = O.Conf(2., (6, 6, 6), mesh_shape=2, float_dtype=np.float64,
                  a_start=1 / 4, a_stop=1., a_nbody_maxstep=1 / 8)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    pmid, disp, vel, _ = O.gen_grid(conf, vel=True)
    rng = np.random.default_rng(3)
    disp = disp + 0.5 * rng.standard_normal(disp.shape)
    vel = 0.2 * rng.standard_normal(disp.shape)
    fwd = O.nbody(dict(pmid=pmid, disp=disp, vel=vel), cosmo, conf)
    back = O.nbody(fwd, cosmo, conf, reverse=True)
    np.testing.assert_allclose(back['disp'], disp, atol=1e-10)
    np.testing.assert_allclose(back['vel'], vel, atol=1e-10)


def test_full_model_gradient_chain():
    """grads.py model in the float64 oracle: adjoint chain (scatter_adj -> nbody_adj ->
    lpt_vjp_modes) vs a central finite difference along a random white-noise direction."""
    n = 8
    kw = dict(a_start=1 / 16, a_nbody_maxstep=1 / 4)
    conf = O.Conf(1., (n,) * 3, mesh_shape=2, float_dtype=np.float64, **kw)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    white = np.random.default_rng(1).standard_normal(conf.ptcl_grid_shape)
    target = np.random.default_rng(0).random(conf.mesh_shape) * 2

    def model(w):
        ic = O.lpt(O.linear_modes(w, cosmo, conf), cosmo, conf)
        out = O.nbody(ic, cosmo, conf)
        return O.scatter(out['pmid'], out['disp'], conf), out

    dens, out = model(white)
    r = dens - target
    dens_cot = 2 * (r - r.mean()) / r.size
    dcot, _ = O.scatter_adj(out['pmid'], out['disp'], conf, dens_cot)
    cot = dict(disp=dcot, vel=np.zeros_like(dcot), acc=np.zeros_like(dcot))
    _, pc, _ = O.nbody_adj(out, cot, cosmo, conf)
    g = O.lpt_vjp_modes(white, cosmo, conf, pc['disp'], pc['vel'])
    v = np.random.default_rng(2).standard_normal(white.shape)
    f = lambda w: np.var(model(w)[0] - target)
    # small h: CIC is only piecewise smooth, larger steps cross cell boundaries
    np.testing.assert_allclose(np.sum(g * v), _fd(f, white, v, h=1e-6), rtol=1e-7)
