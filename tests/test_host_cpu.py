"""CPU tests of the host-side (float64, torch) code that feeds the kernels: configuration
validation, particle grid construction, growth / transfer tables and the kick/drift factors with
their cosmology gradients, against the NumPy/scipy oracle.  No GPU, no kernels."""
import numpy as np
import pytest
import torch

import oracle as O
import pmwd_b200 as pm
from pmwd_b200._lib import PmwdError


def _confs(shape=(4, 4, 4), **kw):
    return (pm.Configuration(1., shape, mesh_shape=2, device='cpu', **kw),
            O.Conf(1., shape, mesh_shape=2, **kw))


def test_configuration_validation_and_properties():
    """configuration.py:148-196 (errors) and :198-320 (derived properties)."""
    conf, oconf = _confs((4, 6, 8))
    assert conf.mesh_shape == oconf.mesh_shape == (8, 12, 16)
    assert conf.cell_size == oconf.cell_size and conf.ptcl_num == oconf.ptcl_num
    assert conf.a_nbody_num == oconf.a_nbody_num == 63
    np.testing.assert_array_equal(conf.a_nbody.numpy(), oconf.a_nbody)
    np.testing.assert_array_equal(conf.growth_a.numpy(), oconf.growth_a)
    np.testing.assert_allclose(conf.transfer_k.numpy(), oconf.transfer_k, rtol=0)
    with pytest.raises(ValueError):
        pm.Configuration(1., (4, 4, 4), mesh_shape=(8, 8), device='cpu')
    with pytest.raises(ValueError):
        pm.Configuration(1., (4, 4, 4), mesh_shape=(2, 2, 2), device='cpu')
    with pytest.raises(ValueError):
        pm.Configuration(1., (4, 4, 4), mesh_shape=(8, 8, 12), device='cpu')
    with pytest.raises(ValueError):
        pm.Configuration(1., (4, 4, 4), symp_splits=((0, 0.5), (1, 0.4)), device='cpu')
    with pytest.raises(ValueError):
        pm.Configuration(1., (4, 4, 4), float_dtype=torch.float64, device='cpu')
    c2 = conf.replace(a_nbody_maxstep=0.1)
    assert c2.a_nbody_num == 10 and conf.a_nbody_num == 63       # frozen, replace() copies


def test_gen_grid_and_from_pos_match_oracle():
    """particles.py:81-144: same pmid / disp as the oracle, incl. a non-integer mesh ratio."""
    for shape, mesh in (((4, 4, 4), 2), ((4, 6, 8), 1.5)):
        conf = pm.Configuration(1.3, shape, mesh_shape=mesh, device='cpu')
        oconf = O.Conf(1.3, shape, mesh_shape=mesh)
        p = pm.Particles.gen_grid(conf, vel=True)
        opmid, odisp, _, _ = O.gen_grid(oconf)
        assert p.pmid.dtype == torch.int16
        np.testing.assert_array_equal(p.pmid.numpy(), opmid)
        np.testing.assert_array_equal(p.disp.numpy(), odisp)
        assert torch.all(p.vel == 0) and p.acc is None
        pos = p.pos()
        q = pm.Particles.from_pos(conf, pos)
        np.testing.assert_allclose(q.pos().numpy(), pos.numpy(), atol=1e-6)


def test_growth_and_transfer_tables_vs_oracle():
    """boltzmann.py:32-229: Dopri5 @ sqrt(eps) (as the reference) vs scipy DOP853 @ 1e-11."""
    conf, oconf = _confs()
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    oc = O.boltzmann(O.SimpleLCDM(oconf), oconf)
    assert np.abs(cosmo.growth.numpy() - oc.growth).max() < 1e-7
    np.testing.assert_allclose(cosmo.transfer.numpy(), oc.transfer, rtol=1e-12)
    k = np.array([0., 1e-3, 0.05, 0.7, 3.], dtype=np.float32)
    np.testing.assert_allclose(pm.linear_power(torch.from_numpy(k), None, cosmo, conf).numpy(),
                               O.linear_power(k, None, oc, oconf), rtol=2e-6)
    for a in (1 / 64, 0.3, 1.0):
        np.testing.assert_allclose(float(pm.E2(a, cosmo)), O.E2(a, oc), rtol=1e-14)
        np.testing.assert_allclose(float(pm.H_deriv(a, cosmo)), O.H_deriv(a, oc), rtol=1e-12)
        np.testing.assert_allclose(float(pm.Omega_m_a(a, cosmo)), O.Omega_m_a(a, oc), rtol=1e-14)


def test_step_factors_and_gradients_vs_oracle():
    """nbody.py:12-36 in float64: torch host code vs the NumPy/scipy oracle, and the autograd
    factor gradients (replacing jax.value_and_grad, nbody.py:51-52,82-83) vs the oracle's
    finite differences."""
    from pmwd_b200.nbody import kick_factor, drift_factor, _factor_valgrad
    conf, oconf = _confs()
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    ocosmo = O.boltzmann(O.SimpleLCDM(oconf), oconf)
    same = ocosmo.replace(growth=cosmo.growth.numpy())            # identical table
    a = oconf.a_nbody
    for i in (0, 10, 40, 62):
        a0, a1 = a[i], a[i + 1]
        am = 0.5 * (a0 + a1)
        for fun, ofun, args in ((kick_factor, O.kick_factor, (a0, a0, am)),
                                (drift_factor, O.drift_factor, (am, a0, a1)),
                                (kick_factor, O.kick_factor, (a1, am, a1))):
            val, grads = _factor_valgrad(fun, *args, cosmo, conf)
            np.testing.assert_allclose(val, ofun(*args, same, oconf), rtol=1e-12)
            np.testing.assert_allclose(val, ofun(*args, ocosmo, oconf), rtol=1e-5)
            _, og = O.factor_grads(ofun, *args, same, oconf)
            np.testing.assert_allclose(grads['Omega_m'].item(), og['Omega_m'], rtol=2e-5, atol=1e-9)
            np.testing.assert_allclose(grads['growth'].numpy(), og['growth'], rtol=1e-5, atol=1e-9)


def test_growth_table_gradient_vs_finite_difference():
    """d(growth table)/d(Omega_m) by autograd through the Runge-Kutta steps vs central
    differences of the oracle's independent ODE solve."""
    conf, oconf = _confs()
    Om = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf, Omega_m=Om), conf)
    w = torch.from_numpy(np.random.default_rng(0).standard_normal(cosmo.growth.shape))
    (g,) = torch.autograd.grad((cosmo.growth * w).sum(), Om)
    h = 1e-5
    gp = O.boltzmann(O.SimpleLCDM(oconf, Omega_m=0.3 + h), oconf).growth
    gm = O.boltzmann(O.SimpleLCDM(oconf, Omega_m=0.3 - h), oconf).growth
    fd = ((gp - gm) / (2 * h) * w.numpy()).sum()
    np.testing.assert_allclose(g.item(), fd, rtol=1e-4)


def test_no_cpu_path_for_kernels():
    """CPU tensors are rejected by every kernel-backed API (there is no CPU fallback)."""
    conf, _ = _confs()
    ptcl = pm.Particles.gen_grid(conf, vel=True)
    cosmo = pm.SimpleLCDM(conf)
    for call in (lambda: pm.scatter(ptcl, conf),
                 lambda: pm.gather(ptcl, conf, torch.zeros(conf.mesh_shape)),
                 lambda: pm.gravity(1., ptcl, cosmo, conf),
                 lambda: pm.nbody(ptcl, None, cosmo, conf),
                 lambda: pm.lpt(torch.zeros((4, 4, 3), dtype=torch.complex64), cosmo, conf),
                 lambda: pm.powspec(torch.zeros(conf.mesh_shape), 1.0),
                 lambda: pm.nbody_step_host(0.5, 0.6, dict(pmid=ptcl.pmid, disp=ptcl.disp, vel=ptcl.vel,
                                                           acc=ptcl.vel), cosmo, conf)):
        with pytest.raises(PmwdError):
            call()


def test_lpt_source_reuses_strain_transforms(monkeypatch):
    """``lpt._L`` (``pmwd/lpt.py:40-76``): the same terms in the same order as the reference's loops,
    each strain component transformed once (6 instead of 9 inverse FFTs at 2LPT, 12 instead of 15 for
    two potentials).  Checked structurally with a stand-in for the strain transform."""
    import types
    import sys
    import pmwd_b200.lpt  # noqa: F401
    mod = sys.modules['pmwd_b200.lpt']
    calls = []

    def fake(kvec, i, j, pot, conf):
        tag = int(pot.sum().item())
        calls.append((tag, i, j))
        g = torch.Generator().manual_seed(100 * tag + 10 * i + j)
        return torch.randn(4, 4, 4, generator=g)

    monkeypatch.setattr(mod, '_strain', fake)
    conf = types.SimpleNamespace(dim=3, ptcl_grid_shape=(4, 4, 4), float_dtype=torch.float32)

    def reference_loops(pot_m, pot_n):
        same = pot_n is None
        if same:
            pot_n = pot_m
        out = torch.zeros(4, 4, 4)
        for i in range(3):
            sm = fake(None, i, i, pot_m, conf)
            for j in range(2, i, -1):
                out = out + sm * fake(None, j, j, pot_n, conf)
            if not same:
                for j in range(i - 1, -1, -1):
                    out = out + sm * fake(None, j, j, pot_n, conf)
        if not same:
            out = out * 0.5
        for i in range(2):
            for j in range(i + 1, 3):
                sm = fake(None, i, j, pot_m, conf)
                sn = sm if same else fake(None, j, i, pot_n, conf)
                out = out - sm * sn
        return out

    p1, p2 = torch.ones(1), torch.full((1,), 2.)
    for pots, n_ours, n_ref in (((p1, None), 6, 9), ((p1, p2), 12, 15)):
        calls.clear()
        got = mod._L(None, *pots, conf)
        assert len(calls) == n_ours and len(set(calls)) == n_ours
        calls.clear()
        want = reference_loops(*pots)
        assert len(calls) == n_ref
        assert torch.equal(got, want)


@pytest.mark.parametrize('spacing, mesh_shape', [(1.0, 2), (1.0, 3), (0.1, 2), (0.7, 2), (0.041, 1), (1.0, 2.5)])
def test_slab_descriptor_selects_the_fast_path_for_any_cell_size(spacing, mesh_shape):
    """dist.SlabForce._desc must qualify for the fused slab kernels (pmwd_cic_fast_path >= 1) also when
    conf.cell_size is not float32-representable (1/3, 0.05, 0.35, ...): the x offset is built from the
    float32 cell size enmesh divides by (pm_util.py:120), so divmod leaves no remainder."""
    import ctypes as C
    import types
    import torch
    from pmwd_b200 import _lib
    from pmwd_b200.configuration import Configuration
    from pmwd_b200.dist import SlabForce
    conf = Configuration(spacing, (8, 8, 8), mesh_shape=mesh_shape, device='cpu')
    Mx = conf.mesh_shape[0]
    pmid = torch.zeros((4, 3), dtype=torch.int16)
    for P in (2, 4):
        if Mx % P:
            continue
        for rank in range(P):
            comm = types.SimpleNamespace(x0=rank * (Mx // P), mx=Mx // P, rank=rank, size=P)
            for h in (1, 2, 5):
                desc = SlabForce(conf, comm)._desc(pmid, h)
                assert _lib.lib().pmwd_cic_fast_path(C.byref(desc)) == 1, (conf.cell_size, rank, h)
    full = SlabForce(conf, types.SimpleNamespace(x0=0, mx=Mx, rank=0, size=1))._desc(pmid, 0)
    assert _lib.lib().pmwd_cic_fast_path(C.byref(full)) == 2


def test_resort_prediction_uses_the_coming_drift():
    """The storage re-sort's keys come from disp + vel * predict: after the pipelined step i (whose gather pass has
    already applied drift i+1) the forces until the next re-sort see the positions now and one / two drifts later,
    so predict = (reorder_every - 1) / 2 x drift_factor of step i+2 (nbody.py:12-22), and 0 when no step follows,
    when the step was not pipelined, or with reorder_every < 2.  Host logic only (no kernels are launched)."""
    from pmwd_b200.nbody import _Stepper, drift_factor, _f32

    conf = pm.Configuration(1., (4, 4, 4), mesh_shape=2, device='cpu', a_nbody_maxstep=0.2)
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    a = conf.a_nbody.tolist()

    class _FakeStore:
        ptcl = pm.Particles.gen_grid(conf)

    st = _Stepper(_FakeStore(), a, cosmo, conf)
    st.fused = True
    n = st.nsteps
    assert n >= 4
    for i_done in range(1, n + 1):            # state right after step i_done - 1: self.i == i_done
        st.i, st.pre = i_done, i_done < n
        got = st.predict()
        if i_done + 1 >= n:
            assert got == 0.0
        else:
            j = i_done + 1                    # the step whose drift takes x_{i_done} to x_{i_done + 1}
            am = a[j] * 0.5 + a[j + 1] * 0.5
            want = _f32(drift_factor(am, a[j], a[j + 1], cosmo, conf))
            assert got == pytest.approx(0.5 * (conf.reorder_every - 1) * want, rel=1e-12) and got > 0
    st.i, st.pre = 1, False                   # not pipelined: the positions at sort time are the next force's
    assert st.predict() == 0.0
    st2 = _Stepper(_FakeStore(), a, cosmo, conf.replace(reorder_every=1))
    st2.i, st2.pre = 1, True
    assert st2.predict() == 0.0
