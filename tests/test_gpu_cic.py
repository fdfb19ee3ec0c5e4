"""GPU parity of the CIC kernels (through the C ABI via the Python mirror of the
reference API) against the oracle.  Tolerances: gather / VJP gathers are evaluated in the
oracle's float32 operation order -> bit-exact; scatter (atomic adds in arbitrary order) is
compared per cell with rel. err <= 1e-5 (BASELINE.json) and must be bitwise reproducible
in deterministic mode."""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


def _pm():
    import pmwd_b200
    return pmwd_b200


def _confs(ptcl_grid_shape, mesh_shape=2, spacing=1., **kw):
    pm = _pm()
    conf = pm.Configuration(spacing, ptcl_grid_shape, mesh_shape=mesh_shape, **kw)
    okw = {k: v for k, v in kw.items() if k in ('a_start', 'a_stop', 'a_nbody_maxstep', 'lpt_order',
                                                 'chunk_size', 'symp_splits')}
    oconf = O.Conf(spacing, ptcl_grid_shape, mesh_shape=mesh_shape, **okw)
    return conf, oconf


def _ptcl(conf, oconf, disp_std, seed=0, pmid_dtype=np.int16):
    pmid, disp, _, _ = O.gen_grid(oconf)
    rng = np.random.default_rng(seed)
    disp = (disp + disp_std * rng.standard_normal(disp.shape)).astype(np.float32)
    pm = _pm()
    ptcl = pm.Particles(conf, torch.from_numpy(pmid.astype(pmid_dtype)).cuda(),
                        torch.from_numpy(disp).cuda())
    return pmid, disp, ptcl


def _rel_err(a, b, floor):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


# ------------------------------------------------------------------ reference known answers
@pytest.mark.parametrize('ptcl_num, pos, chan_shape', [
    (3, (-1.,), (2, 1)),
    (5, (1., -3.), (1, 2, 3)),
    (7, (-3., 5., 7.), None),
    (7, (3., -5., 7.), ()),
    (7, (3., 5., -7.), (1,)),
], ids=['1d', '2d', '3d1', '3d2', '3d3'])
@pytest.mark.parametrize('general', [False, True], ids=['fast', 'f64branch'])
def test_scatter_centered_ptcl(ptcl_num, pos, chan_shape, general):
    """/root/reference/tests/pm_test.py:43-71 through the CUDA path: exact n * 2^-dim."""
    pm = _pm()
    dim = len(pos)
    mesh_shape = (2,) * dim
    conf = pm.Configuration(2., (2,) * dim, mesh_shape=mesh_shape)
    assert conf.cell_size == 2.
    pmid = torch.zeros((ptcl_num, dim), dtype=torch.int16, device='cuda')
    disp = torch.tensor(pos, dtype=torch.float32, device='cuda').repeat(ptcl_num, 1)
    ptcl = pm.Particles(conf, pmid, disp)
    val = 1.
    if chan_shape is None:
        chan_shape = ()
    else:
        val = torch.ones((ptcl_num,) + chan_shape, device='cuda')
    mesh = torch.zeros(mesh_shape + chan_shape, device='cuda')
    out = pm.scatter(ptcl, conf, mesh=mesh, val=val, cell_size=2. if general else None)
    assert out.shape == mesh.shape
    assert torch.all(mesh == 0)                                # input untouched
    expect = torch.full_like(out, ptcl_num * 2. ** -dim)
    assert torch.equal(out, expect)


@pytest.mark.parametrize('ptcl_num, dim, chan_shape', [
    (3, 1, (2, 1)), (5, 2, (1, 2, 3)), (7, 3, None), (7, 3, ()), (7, 3, (1,)),
], ids=['1d', '2d', '3d1', '3d2', '3d3'])
def test_scatter_sum_gather_uniform(ptcl_num, dim, chan_shape):
    """/root/reference/tests/pm_test.py:86-118: mass conservation; gather of ones = ones."""
    pm = _pm()
    rng = np.random.default_rng(0)
    for mesh_n, kind in ((3, 'scatter'), (5, 'gather')):
        conf = pm.Configuration(2., (mesh_n,) * dim, mesh_shape=1)
        pmid = torch.zeros((ptcl_num, dim), dtype=torch.int16, device='cuda')
        disp = torch.from_numpy((7. * rng.standard_normal((ptcl_num, dim))).astype(np.float32)).cuda()
        ptcl = pm.Particles(conf, pmid, disp)
        cs = () if chan_shape is None else chan_shape
        if kind == 'scatter':
            val = 1. if chan_shape is None else torch.ones((ptcl_num,) + cs, device='cuda')
            out = pm.scatter(ptcl, conf, mesh=torch.zeros((mesh_n,) * dim + cs, device='cuda'), val=val)
            np.testing.assert_allclose(out.sum().item(), ptcl_num * np.prod(cs), rtol=1e-6)
        else:
            val = 0. if chan_shape is None else torch.zeros((ptcl_num,) + cs, device='cuda')
            out = pm.gather(ptcl, conf, torch.ones((mesh_n,) * dim + cs, device='cuda'), val=val)
            np.testing.assert_allclose(out.cpu().numpy(), 1., rtol=0, atol=2e-7)


def test_channel_mismatch_raises():
    """scatter.py:45-47 / gather.py:41-43: ValueError on channel-shape mismatch."""
    pm = _pm()
    conf, oconf = _confs((2, 2, 2))
    _, _, ptcl = _ptcl(conf, oconf, 0.1)
    with pytest.raises(ValueError):
        pm.scatter(ptcl, conf, mesh=torch.zeros(4, 4, 4, 2, device='cuda'), val=torch.ones(8, 3, device='cuda'))
    with pytest.raises(ValueError):
        pm.gather(ptcl, conf, torch.zeros(4, 4, 4, 2, device='cuda'))


def test_cpu_tensors_rejected():
    """No CPU fallback: CPU tensors raise instead of being processed."""
    pm = _pm()
    from pmwd_b200._lib import PmwdError
    conf = pm.Configuration(1., (2, 2, 2), mesh_shape=2)
    ptcl = pm.Particles(conf, torch.zeros(8, 3, dtype=torch.int16), torch.zeros(8, 3))
    with pytest.raises(PmwdError):
        pm.scatter(ptcl, conf)


# ------------------------------------------------------------------ fast path vs oracle
@pytest.mark.parametrize('n, disp_std', [(16, 0.3), (32, 3.0), (64, 8.0)])
@pytest.mark.parametrize('mode', ['atomic', 'deterministic'])
def test_scatter_density_vs_oracle(n, disp_std, mode):
    pm = _pm()
    conf, oconf = _confs((n, n, n), scatter_mode=mode)
    pmid, disp, ptcl = _ptcl(conf, oconf, disp_std)
    ref = O.scatter(pmid, disp, oconf)
    out = pm.scatter(ptcl, conf)
    got = out.cpu().numpy()
    assert got.shape == ref.shape
    # per-cell density rel. err <= 1e-5 (cells with rho ~ 0: abs tol 1e-5 * mean, mean = 1)
    assert _rel_err(got, ref, 1.0).max() <= 1e-5
    np.testing.assert_allclose(got.sum(dtype=np.float64), conf.mesh_size, rtol=1e-6)
    if mode == 'deterministic':
        again = pm.scatter(ptcl, conf)
        assert torch.equal(out, again)                         # bitwise reproducible
        # and independent of the atomic path's ordering noise
        f64 = O.scatter(pmid, disp.astype(np.float64),
                        O.Conf(1., (n, n, n), mesh_shape=2, float_dtype=np.float64))
        assert _rel_err(got, f64, 1.0).max() <= 1e-5


def test_scatter_odd_mesh_and_int_dtypes():
    """Odd mesh sizes (no aligned z pairs), non-cubic boxes, int8 / int32 pmid."""
    pm = _pm()
    for shape, dt, tdt in (((6, 10, 14), np.int8, torch.int8), ((5, 7, 9), np.int16, torch.int16),
                           ((4, 6, 8), np.int32, torch.int32)):
        conf = pm.Configuration(1.5, shape, mesh_shape=1, pmid_dtype=tdt)
        oconf = O.Conf(1.5, shape, mesh_shape=1, pmid_dtype=dt)
        pmid, disp, ptcl = _ptcl(conf, oconf, 2.0, pmid_dtype=dt)
        ref = O.scatter(pmid, disp, oconf)
        got = pm.scatter(ptcl, conf).cpu().numpy()
        assert _rel_err(got, ref, 1.0).max() <= 1e-5
        mesh = np.random.default_rng(1).standard_normal(oconf.mesh_shape).astype(np.float32)
        ref = O.gather(pmid, disp, oconf, mesh)
        got = pm.gather(ptcl, conf, torch.from_numpy(mesh).cuda()).cpu().numpy()
        np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize('chan_shape', [(), (3,), (2, 2)])
def test_gather_bit_exact_vs_oracle(chan_shape):
    pm = _pm()
    conf, oconf = _confs((16, 16, 16))
    pmid, disp, ptcl = _ptcl(conf, oconf, 2.5)
    rng = np.random.default_rng(3)
    mesh = rng.standard_normal(oconf.mesh_shape + chan_shape).astype(np.float32)
    val = rng.standard_normal((len(pmid),) + chan_shape).astype(np.float32) if chan_shape else 0.5
    ref = O.gather(pmid, disp, oconf, mesh, val=val)
    v = torch.from_numpy(val).cuda() if chan_shape else val
    got = pm.gather(ptcl, conf, torch.from_numpy(mesh).cuda(), val=v).cpu().numpy()
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize('offset, cell_size, mesh_mult', [
    (0, 0.25, 2), ((0.1, -0.3, 0.7), None, 1), (0.37, 0.4, 1), ((1.5, 0.2, -2.1), 0.75, 1)])
def test_general_branch_vs_oracle(offset, cell_size, mesh_mult):
    """offset / cell_size (float64 enmesh branch, pm_util.py:99-118), user mesh of another
    shape with out-of-range neighbours dropped (optim.py:30-33 use case)."""
    pm = _pm()
    conf, oconf = _confs((8, 8, 8))
    pmid, disp, ptcl = _ptcl(conf, oconf, 1.7)
    tshape = tuple(mesh_mult * s for s in oconf.mesh_shape)
    if cell_size == 0.75:
        tshape = (9, 11, 10)      # smaller than the wrapped extent: exercises the drop
    rng = np.random.default_rng(5)
    mesh0 = rng.standard_normal(tshape).astype(np.float32)
    ref = O.scatter(pmid, disp, oconf, mesh=mesh0, val=1., offset=offset, cell_size=cell_size)
    got = pm.scatter(ptcl, conf, mesh=torch.from_numpy(mesh0).cuda(), val=1., offset=offset,
                     cell_size=cell_size).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=2e-6)
    ref = O.gather(pmid, disp, oconf, mesh0, val=0., offset=offset, cell_size=cell_size)
    got = pm.gather(ptcl, conf, torch.from_numpy(mesh0).cuda(), val=0., offset=offset,
                    cell_size=cell_size).cpu().numpy()
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize('shape', [(9,), (4, 9), (7, 8), (6, 5, 4)])
@pytest.mark.parametrize('cell_size', [None, 3.])
def test_vjps_vs_oracle(shape, cell_size):
    """custom VJP parity (pm_test.py:121-151 shapes): disp/val/mesh cotangents of scatter and
    gather through torch autograd vs the oracle's analytic adjoints."""
    pm = _pm()
    dim = len(shape)
    conf = pm.Configuration(1.3, shape, mesh_shape=1)
    oconf = O.Conf(1.3, shape, mesh_shape=1)
    pmid, disp, ptcl = _ptcl(conf, oconf, 3.0)
    chan = (2, 1)
    rng = np.random.default_rng(7)
    val = (1 + rng.standard_normal((len(pmid),) + chan)).astype(np.float32)
    mesh = rng.standard_normal(oconf.mesh_shape + chan).astype(np.float32)
    mesh_cot = rng.standard_normal(mesh.shape).astype(np.float32)
    val_cot = rng.standard_normal(val.shape).astype(np.float32)

    d = ptcl.disp.clone().requires_grad_(True)
    v = torch.from_numpy(val).cuda().requires_grad_(True)
    m = torch.from_numpy(mesh).cuda().requires_grad_(True)
    p = ptcl.replace(disp=d)
    out = pm.scatter(p, conf, mesh=m, val=v, cell_size=cell_size)
    out.backward(torch.from_numpy(mesh_cot).cuda())
    dc_ref, vc_ref = O.scatter_adj(pmid, disp, oconf, mesh_cot, val=val, cell_size=cell_size)
    np.testing.assert_array_equal(d.grad.cpu().numpy(), dc_ref)
    np.testing.assert_array_equal(v.grad.cpu().numpy(), vc_ref)
    np.testing.assert_array_equal(m.grad.cpu().numpy(), mesh_cot)

    d.grad = None; v.grad = None; m.grad = None
    out = pm.gather(p, conf, m, val=v, cell_size=cell_size)
    out.backward(torch.from_numpy(val_cot).cuda())
    dc_ref, mc_ref = O.gather_adj(pmid, disp, oconf, mesh, val_cot, cell_size=cell_size)
    np.testing.assert_array_equal(d.grad.cpu().numpy(), dc_ref)
    np.testing.assert_allclose(m.grad.cpu().numpy(), mc_ref, rtol=2e-6, atol=2e-6)
    np.testing.assert_array_equal(v.grad.cpu().numpy(), val_cot)


def test_empty_particles():
    pm = _pm()
    conf = pm.Configuration(1., (2, 2, 2), mesh_shape=2)
    ptcl = pm.Particles(conf, torch.zeros((0, 3), dtype=torch.int16, device='cuda'),
                        torch.zeros((0, 3), device='cuda'))
    out = pm.scatter(ptcl, conf)
    assert out.shape == (4, 4, 4) and torch.all(out == 0)
    g = pm.gather(ptcl, conf, torch.ones(4, 4, 4, device='cuda'))
    assert g.shape == (0,)


# ------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize('n', [256])
@pytest.mark.parametrize('mode', ['atomic', 'deterministic'])
def test_fullsize_mass_conservation_and_uniform_gather(n, mode):
    """BASELINE config sizes (256^3 particles, 512^3 mesh) through size-independent
    properties: total mass = N_m; gather(ones) = 1; linearity of gather."""
    pm = _pm()
    conf = pm.Configuration(1., (n, n, n), mesh_shape=2, scatter_mode=mode)
    ptcl = pm.Particles.gen_grid(conf)
    g = torch.Generator(device='cuda').manual_seed(0)
    disp = ptcl.disp + 4.0 * torch.randn(ptcl.disp.shape, device='cuda', generator=g)
    ptcl = ptcl.replace(disp=disp)
    dens = pm.scatter(ptcl, conf)
    np.testing.assert_allclose(dens.double().sum().item(), conf.mesh_size, rtol=1e-7)
    assert dens.min().item() >= 0.
    ones = torch.ones(conf.mesh_shape, device='cuda')
    out = pm.gather(ptcl, conf, ones)
    assert (out - 1).abs().max().item() <= 3e-7
    # <gather(m), v> == <m, scatter(v)>   (scatter and gather are each other's transpose)
    m = torch.randn(conf.mesh_shape, device='cuda', generator=g)
    v = torch.randn(conf.ptcl_num, device='cuda', generator=g)
    lhs = (pm.gather(ptcl, conf, m).double() * v.double()).sum().item()
    rhs = (pm.scatter(ptcl, conf, mesh=torch.zeros_like(m), val=v).double() * m.double()).sum().item()
    np.testing.assert_allclose(lhs, rhs, rtol=1e-5)
