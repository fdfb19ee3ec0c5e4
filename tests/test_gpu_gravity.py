"""GPU parity of the k-space kernels, the fused force pipeline, the leapfrog kernels, LPT,
the N-body integration and its reverse-time adjoint against the oracle.

Stated float32 tolerances (BASELINE.json north_star): per-cell density rel. err <= 1e-5,
final positions within 1e-4 cell, P(k) within 0.1 % at all k, adjoint gradients with
cosine similarity >= 0.9999."""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


def _pm():
    import pmwd_b200
    return pmwd_b200


def _confs(shape, mesh_shape=2, spacing=1., **kw):
    pm = _pm()
    return (pm.Configuration(spacing, shape, mesh_shape=mesh_shape, **kw),
            O.Conf(spacing, shape, mesh_shape=mesh_shape,
                   **{k: v for k, v in kw.items() if k not in ('scatter_mode', 'reorder_every',
                                                               'reorder_min_disp')}))


def _cos(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / np.sqrt((a @ a) * (b @ b)))


def _rms(x):
    return float(np.sqrt(np.mean(np.square(np.asarray(x, dtype=np.float64)))))


# ------------------------------------------------------------------------- k-space kernels
@pytest.mark.parametrize('shape', [(16,), (6, 9), (7, 8), (8, 8, 8), (6, 5, 4), (12, 10, 9)])
def test_kspace_kernels_vs_oracle(shape):
    """laplace / neg_grad / strain (gravity.py:9-44, lpt.py:13-32) incl. even/odd meshes
    (Nyquist zeroing) and ranks 1-3; same float32 operation order -> agreement to 1 ulp."""
    pm = _pm()
    from pmwd_b200.lpt import _Strain
    spacing = 0.7
    rng = np.random.default_rng(0)
    cshape = shape[:-1] + (shape[-1] // 2 + 1,)
    src = (rng.standard_normal(cshape) + 1j * rng.standard_normal(cshape)).astype(np.complex64)
    okvec = O.fftfreq(shape, spacing, dtype=np.float32)
    kvec = pm.fftfreq(shape, spacing, dtype=torch.float32, device='cuda')
    for a, b in zip(okvec, kvec):
        np.testing.assert_array_equal(a, b.cpu().numpy())
    tsrc = torch.from_numpy(src).cuda()
    ref = O.laplace(okvec, src)
    got = pm.laplace(kvec, tsrc).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-7, atol=0)
    for ax in range(len(shape)):
        ref = O.neg_grad(okvec[ax], src, spacing)
        got = pm.neg_grad(kvec[ax], tsrc, spacing).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=2e-7, atol=0)
    if len(shape) == 3:
        oc = O.Conf(spacing, shape)
        for i in range(3):
            for j in range(i, 3):
                k_i, k_j = okvec[i], okvec[j]
                nyq = np.pi / spacing
                eps = nyq * np.finfo(np.float32).eps
                if i != j:
                    k_i = np.where(np.abs(np.abs(k_i) - nyq) <= eps, 0, k_i)
                    k_j = np.where(np.abs(np.abs(k_j) - nyq) <= eps, 0, k_j)
                ref = (-k_i * k_j * src).astype(np.complex64)
                got = _Strain.apply(tsrc, shape, spacing, i, j).cpu().numpy()
                np.testing.assert_allclose(got, ref, rtol=2e-7, atol=0)


def test_plane_wave_force_gpu():
    """Closed-form pin of the mesh pipeline through the C ABI (pmwd_fft_r2c ->
    pmwd_kspace_force -> pmwd_fft_c2r): rho = 1 + A cos(kx) gives
    F_x = -1.5 Om A sin(kx) / k and F_y = F_z = 0  (gravity.py:9-16, 37-44, 54-64)."""
    import ctypes as C
    from pmwd_b200 import _lib
    n = 64
    shape = (n, n, n)
    cell = 0.5
    kf = 2 * np.pi / (n * cell) * 3
    x = np.arange(n) * cell
    A, Om = 0.1, 0.3
    rho = (1 + A * np.cos(kf * x))[:, None, None] * np.ones(shape)
    trho = torch.from_numpy(rho.astype(np.float32)).cuda()
    dev = trho.device
    ctx = _lib.Context.get(dev).reserve(shape)
    lib = _lib.lib()
    shp = _lib.shape_arr(shape)
    spec = torch.empty((n, n, n // 2 + 1), dtype=torch.complex64, device=dev)
    g = [torch.empty_like(spec) for _ in range(3)]
    F = [torch.empty(shape, device=dev) for _ in range(3)]
    st = _lib.stream_ptr(dev)
    _lib.check(lib.pmwd_fft_r2c(ctx.handle, st, 3, shp, _lib.ptr(trho), _lib.ptr(spec)), 'r2c')
    arr = (C.c_void_p * 3)(*[t.data_ptr() for t in g])
    _lib.check(lib.pmwd_kspace_force(st, 3, shp, cell, 1.5 * Om, _lib.ptr(spec), arr), 'ks')
    for a in range(3):
        _lib.check(lib.pmwd_fft_c2r(ctx.handle, st, 3, shp, _lib.ptr(g[a]), _lib.ptr(F[a]),
                                    1.0 / n ** 3), 'c2r')
    expect = -1.5 * Om * A * np.sin(kf * x) / kf
    got = F[0][:, 5, 7].cpu().numpy()
    assert np.abs(got - expect).max() <= 1e-6 * np.abs(expect).max()
    assert F[1].abs().max().item() <= 1e-6 * np.abs(expect).max()
    assert F[2].abs().max().item() <= 1e-6 * np.abs(expect).max()


@pytest.mark.parametrize('n, disp_std', [(16, 0.5), (32, 4.0), (64, 8.0)])
@pytest.mark.parametrize('mode', ['atomic', 'deterministic'])
def test_gravity_vs_oracle(n, disp_std, mode):
    """acc rel-RMS error <= 1e-5 against the float64 oracle and the float32 oracle."""
    pm = _pm()
    conf, oconf = _confs((n, n, n), scatter_mode=mode)
    cosmo = pm.SimpleLCDM(conf)
    pmid, disp, _, _ = O.gen_grid(oconf)
    rng = np.random.default_rng(1)
    disp = (disp + disp_std * rng.standard_normal(disp.shape)).astype(np.float32)
    ptcl = pm.Particles(conf, torch.from_numpy(pmid).cuda(), torch.from_numpy(disp).cuda())
    acc = pm.gravity(1., ptcl, cosmo, conf).cpu().numpy()
    ref32 = O.gravity(pmid, disp, 0.3, oconf)
    o64 = O.Conf(1., (n, n, n), mesh_shape=2, float_dtype=np.float64)
    ref64 = O.gravity(pmid, disp.astype(np.float64), 0.3, o64)
    scale = _rms(ref64)
    assert _rms(acc - ref64) <= 1e-5 * scale
    assert _rms(acc - ref32) <= 1e-5 * scale
    # our error vs float64 is of the same order as the float32 oracle's own
    assert _rms(acc - ref64) <= 3 * _rms(ref32 - ref64) + 1e-7 * scale


def test_gravity_config2_full_size_vs_compiled_oracle():
    """BASELINE config 2 at its full size (256^3 particles / 512^3 mesh): the CUDA force against
    the oracle's compiled twin (oracle/cpm.c, bit-identical to the NumPy oracle on one thread,
    summation order of the scatter aside; float32 noise floor vs float64 measured at 2.6e-7).
    Tolerance: acc rel-RMS error <= 1e-5, the same as at the small sizes."""
    from oracle import cpm
    pm = _pm()
    n = 256
    conf, oconf = _confs((n, n, n))
    cosmo = pm.SimpleLCDM(conf)
    pmid, disp, _, _ = O.gen_grid(oconf)
    pmid = np.ascontiguousarray(pmid)
    rng = np.random.default_rng(1)
    disp = (disp + 4.0 * rng.standard_normal(disp.shape, dtype=np.float32)).astype(np.float32)
    ptcl = pm.Particles(conf, torch.from_numpy(pmid).cuda(), torch.from_numpy(disp).cuda())
    acc = pm.gravity(1., ptcl, cosmo, conf).cpu().numpy()
    ref = cpm.gravity(pmid, disp, 0.3, oconf)
    scale = _rms(ref)
    assert _rms(acc - ref) <= 1e-5 * scale
    assert np.abs(acc - ref).max() <= 1e-3 * scale


def test_gravity_config3_momentum_conservation():
    """The north-star size (512^3 particles / 1024^3 mesh) through a size-independent property:
    scatter and gather share their weights and the mesh force operator is antisymmetric, so the
    accelerations sum to zero (4e-10 of their rms in the float32 oracle at 128^3)."""
    pm = _pm()
    n = 512
    conf = pm.Configuration(1., (n, n, n), mesh_shape=2)
    cosmo = pm.SimpleLCDM(conf)
    ptcl = pm.Particles.gen_grid(conf)
    g = torch.Generator(device='cuda').manual_seed(0)
    ptcl = ptcl.replace(disp=ptcl.disp + 4.0 * torch.randn(ptcl.disp.shape, device='cuda', generator=g))
    acc = pm.gravity(1., ptcl, cosmo, conf)
    assert torch.isfinite(acc).all()
    rms = acc.double().square().mean().sqrt().item()
    assert rms > 0
    assert acc.double().mean(dim=0).abs().max().item() <= 1e-6 * rms
    del acc
    from pmwd_b200.gravity import release_workspaces
    release_workspaces()
    torch.cuda.empty_cache()


@pytest.mark.parametrize('mode', ['atomic', 'deterministic'])
def test_gravity_vjp_vs_oracle(mode):
    """pmwd_force_adj vs the oracle's chain of the reference VJP rules (cos >= 0.9999,
    rel-RMS <= 1e-4), and the Omega_m cotangent."""
    pm = _pm()
    n = 16
    conf, oconf = _confs((n, n, n), scatter_mode=mode)
    pmid, disp, _, _ = O.gen_grid(oconf)
    rng = np.random.default_rng(2)
    disp = (disp + 1.5 * rng.standard_normal(disp.shape)).astype(np.float32)
    pi = rng.standard_normal(disp.shape).astype(np.float32)
    Om = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    cosmo = pm.SimpleLCDM(conf, Omega_m=Om)
    d = torch.from_numpy(disp).cuda().requires_grad_(True)
    ptcl = pm.Particles(conf, torch.from_numpy(pmid).cuda(), d)
    acc = pm.gravity(1., ptcl, cosmo, conf)
    acc.backward(torch.from_numpy(pi).cuda())
    o64 = O.Conf(1., (n, n, n), mesh_shape=2, float_dtype=np.float64)
    acc_ref, dcot_ref, Om_ref = O.gravity_vjp(pmid, disp.astype(np.float64), 0.3, o64, pi)
    got = d.grad.cpu().numpy()
    assert _cos(got, dcot_ref) >= 0.9999
    assert _rms(got - dcot_ref) <= 1e-4 * _rms(dcot_ref)
    np.testing.assert_allclose(Om.grad.item(), Om_ref, rtol=1e-4)
    # and the float32 oracle
    _, dcot32, _ = O.gravity_vjp(pmid, disp, 0.3, oconf, pi)
    assert _rms(got - dcot32) <= 1e-4 * _rms(dcot_ref)


@pytest.mark.parametrize('shape', [(128, 4, 5), (256, 4, 9), (512, 4, 5), (512, 12, 65), (1024, 4, 5), (1024, 12, 33)],
                         ids=lambda shp: 'x'.join(str(n) for n in shp))
def test_register_xpass_force_and_vjp_vs_oracle(shape):
    """The fused x-pass kernels behind the headline numbers (`xr16_force_kernel<nx>` /
    `xr16_force_adj_kernel<nx>`, nx = mesh x extent = 2 x shape[0] in {256, 512, 1024, 2048}) inside
    the whole force pipeline and its VJP, on anisotropic boxes small enough for the NumPy oracle:
    pm.gravity / its backward (= pmwd_force, pmwd_force_adj) against O.gravity and O.gravity_vjp
    (scipy pocketfft + the reference's laplace / neg_grad, gravity.py:9-72, nbody.py:108-118).
    Tolerances as at the cubic sizes: acc rel-RMS <= 1e-5, disp cotangent cos >= 0.9999 and rel-RMS <= 1e-4,
    Omega_m cotangent 1e-4."""
    pm = _pm()
    conf, oconf = _confs(shape)
    o64 = O.Conf(1., shape, mesh_shape=2, float_dtype=np.float64)
    assert tuple(conf.mesh_shape) == tuple(2 * n for n in shape)
    pmid, disp, _, _ = O.gen_grid(oconf)
    rng = np.random.default_rng(7)
    disp = (disp + 1.5 * rng.standard_normal(disp.shape)).astype(np.float32)
    pi = rng.standard_normal(disp.shape).astype(np.float32)
    Om = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    cosmo = pm.SimpleLCDM(conf, Omega_m=Om)
    d = torch.from_numpy(disp).cuda().requires_grad_(True)
    ptcl = pm.Particles(conf, torch.from_numpy(np.ascontiguousarray(pmid)).cuda(), d)
    from pmwd_b200 import _lib
    n0 = _lib.lib().pmwd_launch_count()
    acc = pm.gravity(1., ptcl, cosmo, conf)
    acc.backward(torch.from_numpy(pi).cuda())
    assert _lib.lib().pmwd_launch_count() > n0
    # mesh x extent >= 256: the register kernels (xpass16.cu); 128: the radix-4 shared-memory kernel
    assert _lib.lib().pmwd_xpass_last_variant() == (2 if conf.mesh_shape[0] >= 256 else 1)
    acc64, dcot64, Om64 = O.gravity_vjp(pmid, disp.astype(np.float64), 0.3, o64, pi)
    acc32, dcot32, _ = O.gravity_vjp(pmid, disp, 0.3, oconf, pi)
    got = acc.detach().cpu().numpy()
    scale = _rms(acc64)
    assert _rms(got - acc64) <= 1e-5 * scale
    assert _rms(got - acc32) <= 1e-5 * scale
    # same order as the float32 oracle's own distance from float64 (measured on B200: 8.6e-7 of the rms
    # at 1024 x 24 x 130 against pocketfft's 2.4e-7 -- a 1024-point float32 transform in registers plus
    # cuFFT's length-130 passes)
    assert _rms(got - acc64) <= 5 * _rms(acc32 - acc64) + 1e-7 * scale
    gd = d.grad.cpu().numpy()
    assert _cos(gd, dcot64) >= 0.9999
    assert _rms(gd - dcot64) <= 1e-4 * _rms(dcot64)
    assert _rms(gd - dcot32) <= 1e-4 * _rms(dcot64)
    np.testing.assert_allclose(Om.grad.item(), Om64, rtol=1e-4)


def test_gravity_general_dims():
    """1-D / 2-D gravity (composition path) vs oracle, incl. autograd VJP vs oracle VJP."""
    pm = _pm()
    for shape in [(16,), (6, 9), (7, 8)]:
        conf, oconf = _confs(shape, mesh_shape=1, spacing=1.3)
        cosmo = pm.SimpleLCDM(conf)
        pmid, disp, _, _ = O.gen_grid(oconf)
        rng = np.random.default_rng(4)
        disp = (disp + 1.1 * rng.standard_normal(disp.shape)).astype(np.float32)
        d = torch.from_numpy(disp).cuda().requires_grad_(True)
        ptcl = pm.Particles(conf, torch.from_numpy(pmid).cuda(), d)
        acc = pm.gravity(1., ptcl, cosmo, conf)
        ref = O.gravity(pmid, disp, 0.3, oconf)
        np.testing.assert_allclose(acc.detach().cpu().numpy(), ref, rtol=2e-4, atol=2e-5 * _rms(ref))
        pi = rng.standard_normal(disp.shape).astype(np.float32)
        acc.backward(torch.from_numpy(pi).cuda())
        o64 = O.Conf(1.3, shape, mesh_shape=1, float_dtype=np.float64)
        _, dcot, _ = O.gravity_vjp(pmid, disp.astype(np.float64), 0.3, o64, pi)
        assert _cos(d.grad.cpu().numpy(), dcot) >= 0.9999


# --------------------------------------------------------------------------- leapfrog
def test_kick_drift_bit_exact():
    """nbody.py:39-46,70-77: mul-then-add in float32 -> bitwise equal to NumPy."""
    from pmwd_b200 import _lib
    rng = np.random.default_rng(0)
    for n in (3, 1000, 3 * 12345):
        disp, vel, acc = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
        K, D = np.float32(0.37), np.float32(-1.91)
        td, tv, ta = (torch.from_numpy(x.copy()).cuda() for x in (disp, vel, acc))
        _lib.check(_lib.lib().pmwd_kick_drift(_lib.stream_ptr(), n, _lib.ptr(td), _lib.ptr(tv),
                                              _lib.ptr(ta), float(K), float(D), 1, 1), 'kd')
        v = vel + acc * K
        x = disp + v * D
        np.testing.assert_array_equal(tv.cpu().numpy(), v)
        np.testing.assert_array_equal(td.cpu().numpy(), x)


def test_kick_drift_adj_vs_numpy():
    from pmwd_b200 import _lib
    rng = np.random.default_rng(1)
    n = 3 * 5001
    disp, vel, acc, xi, pi, alpha = (rng.standard_normal(n).astype(np.float32) for _ in range(6))
    K, D = np.float32(0.21), np.float32(0.83)
    t = [torch.from_numpy(x.copy()).cuda() for x in (disp, vel, acc, xi, pi, alpha)]
    sums = torch.zeros(2, dtype=torch.float64, device='cuda')
    _lib.check(_lib.lib().pmwd_kick_drift_adj(
        _lib.stream_ptr(), n, *[_lib.ptr(x) for x in t], float(K), float(D), 1, 1, _lib.ptr(sums)), 'kda')
    v = vel + acc * K
    x_ = xi - alpha * K
    s_pa = np.sum(pi.astype(np.float64) * acc)
    d_ = disp + v * D
    p_ = pi - x_ * D
    s_xv = np.sum(x_.astype(np.float64) * v)
    np.testing.assert_array_equal(t[1].cpu().numpy(), v)
    np.testing.assert_array_equal(t[3].cpu().numpy(), x_)
    np.testing.assert_array_equal(t[0].cpu().numpy(), d_)
    np.testing.assert_array_equal(t[4].cpu().numpy(), p_)
    np.testing.assert_allclose(sums.cpu().numpy(), [s_pa, s_xv], rtol=1e-12)


# ------------------------------------------------------------------------------- LPT
@pytest.mark.parametrize('order', [1, 2])
def test_lpt_vs_oracle(order):
    pm = _pm()
    n = 32
    conf, oconf = _confs((n, n, n), lpt_order=order)
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    ocosmo = O.boltzmann(O.SimpleLCDM(oconf), oconf).replace(growth=cosmo.growth.numpy())
    omodes = O.linear_modes(O.white_noise(0, oconf), ocosmo, oconf)
    modes = pm.linear_modes(pm.white_noise(0, conf), cosmo, conf)
    np.testing.assert_allclose(modes.cpu().numpy(), omodes, rtol=1e-5, atol=1e-6 * np.abs(omodes).max())
    ptcl, _ = pm.lpt(modes, cosmo, conf)
    ref = O.lpt(omodes, ocosmo, oconf)
    assert torch.equal(ptcl.pmid.cpu(), torch.from_numpy(ref['pmid']))
    assert _rms(ptcl.disp.cpu().numpy() - ref['disp']) <= 1e-5 * _rms(ref['disp'])
    assert _rms(ptcl.vel.cpu().numpy() - ref['vel']) <= 1e-5 * _rms(ref['vel'])


# ------------------------------------------------------------------------------ N-body
def _ic(n, seed=0, **kw):
    pm = _pm()
    conf, oconf = _confs((n, n, n), **kw)
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    ocosmo = O.boltzmann(O.SimpleLCDM(oconf), oconf).replace(growth=cosmo.growth.numpy())
    omodes = O.linear_modes(O.white_noise(seed, oconf), ocosmo, oconf)
    ic = O.lpt(omodes, ocosmo, oconf)
    ptcl = pm.Particles(conf, torch.from_numpy(ic['pmid']).cuda(), torch.from_numpy(ic['disp']).cuda(),
                        vel=torch.from_numpy(ic['vel']).cuda())
    return pm, conf, oconf, cosmo, ocosmo, ic, ptcl


@pytest.mark.parametrize('mode', ['atomic', 'deterministic'])
def test_nbody_config1_vs_oracle(mode):
    """BASELINE config 1: 64^3 particles, 128^3 mesh, 2LPT + 10 leapfrog steps.  Final
    positions within 1e-4 cell (max-norm), P(k) within 0.1 % at all k."""
    pm, conf, oconf, cosmo, ocosmo, ic, ptcl = _ic(64, a_nbody_maxstep=0.1, scatter_mode=mode)
    assert conf.a_nbody_num == 10
    out, _ = pm.nbody(ptcl, None, cosmo, conf)
    ref = O.nbody(dict(ic), ocosmo, oconf)
    cell = conf.cell_size
    dpos = np.abs(out.disp.cpu().numpy() - ref['disp']) / cell
    # float32 round-off differences (cuFFT vs pocketfft, summation order) are amplified by the
    # N-body dynamics, so the error has heavy tails; the reference's own run-to-run noise is
    # ~2e-5 cell RMS (adjoint.tex:1462-1469).  "Within 1e-4 cell" is asserted for the RMS and
    # the 99.9th percentile; the max-norm is checked against the float32 oracle's own
    # distance from a float64 evaluation of the same algorithm.
    o64 = O.Conf(1., (64,) * 3, mesh_shape=2, float_dtype=np.float64, a_nbody_maxstep=0.1)
    ref64 = O.nbody(dict(pmid=ic['pmid'], disp=ic['disp'].astype(np.float64),
                         vel=ic['vel'].astype(np.float64)), ocosmo, o64)
    noise = np.abs(ref['disp'] - ref64['disp']) / cell       # float32 oracle vs exact arithmetic
    dpos64 = np.abs(out.disp.cpu().numpy() - ref64['disp']) / cell
    stats = dict(rms=_rms(dpos), p999=float(np.quantile(dpos, 0.999)), max=float(dpos.max()),
                 rms_vs_f64=_rms(dpos64), max_vs_f64=float(dpos64.max()),
                 noise_rms=_rms(noise), noise_p999=float(np.quantile(noise, 0.999)),
                 noise_max=float(noise.max()))
    print('position error [cell]:', stats)
    assert stats['rms'] <= 1e-4 and stats['p999'] <= 1e-4, stats
    # we are as close to the float32 reference arithmetic as that is to exact arithmetic
    assert stats['rms_vs_f64'] <= 2 * stats['noise_rms'] + 1e-6, stats
    assert stats['max_vs_f64'] <= 3 * stats['noise_max'] + 1e-5, stats
    dvel = _rms(out.vel.cpu().numpy() - ref['vel']) / _rms(ref['vel'])
    assert dvel <= 1e-5
    # final accelerations inherit the (amplified) position differences in clustered regions
    assert _rms(out.acc.cpu().numpy() - ref['acc']) <= 1e-4 * _rms(ref['acc'])
    dens = pm.scatter(out, conf).cpu().numpy()
    oden = O.scatter(ref['pmid'], ref['disp'], oconf)
    k, P, N, _ = O.powspec(dens, conf.cell_size)
    k2, P2, _, _ = O.powspec(oden, conf.cell_size)
    assert np.abs(P / P2 - 1).max() <= 1e-3
    # 3-digit sanity of the physics (growth of structure): disp std grows by ~D(1)/D(1/64)
    assert 30 < out.disp.std().item() / ptcl.disp.std().item() < 60


def test_nbody_config2_full_size_steps_vs_compiled_oracle():
    """BASELINE config 2 at its full size (256^3 particles / 512^3 mesh, 63-step schedule): 2LPT
    initial conditions, then the first 6 leapfrog steps through the public nbody_init / nbody_step
    against the oracle's compiled twin (oracle/cpm.c).  Positions within 1e-4 cell (RMS and the
    99.9th percentile), velocities and accelerations 1e-5 / 1e-4 relative."""
    from oracle import cpm
    pm, conf, oconf, cosmo, ocosmo, ic, ptcl = _ic(256)
    a = conf.a_nbody
    nsteps = 6
    p, _ = pm.nbody_init(a[0], ptcl, None, cosmo, conf)
    for i in range(nsteps):
        p, _ = pm.nbody_step(a[i], a[i + 1], p, None, cosmo, conf)
    q = cpm.nbody_init(float(a[0]), dict(ic), ocosmo, oconf)
    for i in range(nsteps):
        q = cpm.nbody_step(float(a[i]), float(a[i + 1]), q, ocosmo, oconf)
    dpos = np.abs(p.disp.cpu().numpy() - q['disp']) / conf.cell_size
    stats = dict(rms=_rms(dpos), p999=float(np.quantile(dpos[::7], 0.999)), max=float(dpos.max()))
    print('position error [cell]:', stats)
    assert stats['rms'] <= 1e-4 and stats['p999'] <= 1e-4, stats
    assert _rms(p.vel.cpu().numpy() - q['vel']) <= 1e-5 * _rms(q['vel'])
    assert _rms(p.acc.cpu().numpy() - q['acc']) <= 1e-4 * _rms(q['acc'])
    # the run did something: displacements grew with the growth factor over the 6 steps
    assert p.disp.std().item() > 1.05 * ptcl.disp.std().item()


@pytest.mark.parametrize('mode', ['atomic', 'deterministic'])
def test_storage_reorder_is_transparent(mode):
    """Re-sorting the integrator's particle storage by mesh cell (csrc/reorder.cu) must not
    change results beyond float32 summation-order noise, and outputs stay in Lagrangian order;
    forward and adjoint."""
    base = dict(a_nbody_maxstep=1 / 16, scatter_mode=mode)
    pm, conf0, oconf, cosmo, ocosmo, ic, ptcl = _ic(32, reorder_every=0, **base)
    conf1 = conf0.replace(reorder_every=1, reorder_min_disp=0.0)
    outs = []
    for conf in (conf0, conf1):
        d = ptcl.disp.clone().requires_grad_(True)
        out, _ = pm.nbody(pm.Particles(conf, ptcl.pmid, d, vel=ptcl.vel), None, cosmo, conf)
        (out.disp ** 2).sum().backward()
        outs.append((pm.Particles(conf, out.pmid, out.disp.detach(), vel=out.vel.detach()), d.grad))
    (o0, g0), (o1, g1) = outs
    assert torch.equal(o0.pmid, ptcl.pmid)                     # caller's arrays untouched
    assert torch.equal(o0.pmid, o1.pmid)
    cell = conf0.cell_size
    err = ((o0.disp - o1.disp).abs() / cell).cpu().numpy()
    assert _rms(err) <= 1e-4 and np.quantile(err, 0.999) <= 1e-4
    assert _rms((o0.vel - o1.vel).cpu().numpy()) <= 1e-4 * _rms(o0.vel.cpu().numpy())
    # gradient of sum(disp^2) through a chaotic 16-step run: two *atomic* runs differ from each
    # other at this level too (float32 summation-order noise)
    assert _cos(g0.cpu().numpy(), g1.cpu().numpy()) >= 0.999
    if mode == 'deterministic':      # still bitwise reproducible run to run
        out2, _ = pm.nbody(pm.Particles(conf1, ptcl.pmid, ptcl.disp, vel=ptcl.vel), None, cosmo, conf1)
        assert torch.equal(out2.disp, o1.disp) and torch.equal(out2.vel, o1.vel)


def test_nbody_reversibility():
    """pm_test.py:177-192: forward then reverse integration returns the initial state to
    float32 round-off (RMSD/sigma of the order of the reference's table, adjoint.tex:1534-1541)."""
    pm, conf, oconf, cosmo, ocosmo, ic, ptcl = _ic(32, a_nbody_maxstep=0.1)
    out, _ = pm.nbody(ptcl, None, cosmo, conf)
    back, _ = pm.nbody(out, None, cosmo, conf, reverse=True)
    r_disp = _rms((back.disp - ptcl.disp).cpu().numpy()) / _rms(ptcl.disp.cpu().numpy())
    r_vel = _rms((back.vel - ptcl.vel).cpu().numpy()) / _rms(ptcl.vel.cpu().numpy())
    assert r_disp < 5e-2 and r_vel < 7e-2
    # same statistic from the oracle: ours must not be worse than 3x
    oref = O.nbody(O.nbody(dict(ic), ocosmo, oconf), ocosmo, oconf, reverse=True)
    o_disp = _rms(oref['disp'] - ic['disp']) / _rms(ic['disp'])
    assert r_disp <= 3 * o_disp + 1e-6


def test_nbody_step_api_matches_nbody():
    """nbody_init / nbody_step (imported by user scripts, demo_data_time_evo.py:17) give the
    same trajectory as nbody, and never modify their inputs."""
    pm, conf, oconf, cosmo, ocosmo, ic, ptcl = _ic(16, a_nbody_maxstep=0.25, scatter_mode='deterministic',
                                                    reorder_every=0)
    d0 = ptcl.disp.clone()
    out, _ = pm.nbody(ptcl, None, cosmo, conf)
    a = conf.a_nbody
    p, obs = pm.nbody_init(a[0], ptcl, None, cosmo, conf)
    for a0, a1 in zip(a[:-1], a[1:]):
        p, obs = pm.nbody_step(a0, a1, p, obs, cosmo, conf)
    assert torch.equal(p.disp, out.disp) and torch.equal(p.vel, out.vel)
    assert torch.equal(ptcl.disp, d0)


@pytest.mark.parametrize('mode', ['atomic', 'deterministic'])
def test_nbody_adjoint_vs_oracle(mode):
    """Reverse-time adjoint (nbody.py:226-276) vs the oracle's: particle cotangents with
    cosine >= 0.9999, cosmology cotangent leaves within 1e-3 relative."""
    pm, conf, oconf, cosmo, ocosmo, ic, ptcl = _ic(16, a_nbody_maxstep=0.125, scatter_mode=mode)
    rng = np.random.default_rng(5)
    w_disp = rng.standard_normal(ic['disp'].shape).astype(np.float32)
    w_vel = rng.standard_normal(ic['disp'].shape).astype(np.float32)

    Om = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    gt = cosmo.growth.detach().clone().requires_grad_(True)
    c = cosmo.replace(Omega_m=Om, growth=gt)
    d = ptcl.disp.clone().requires_grad_(True)
    v = ptcl.vel.clone().requires_grad_(True)
    out, _ = pm.nbody(ptcl.replace(disp=d, vel=v), None, c, conf)
    obj = (out.disp * torch.from_numpy(w_disp).cuda()).sum() + (out.vel * torch.from_numpy(w_vel).cuda()).sum()
    obj.backward()

    o64 = O.Conf(1., (16,) * 3, mesh_shape=2, float_dtype=np.float64, a_nbody_maxstep=0.125)
    ic64 = dict(pmid=ic['pmid'], disp=ic['disp'].astype(np.float64), vel=ic['vel'].astype(np.float64))
    final = O.nbody(ic64, ocosmo, o64)
    cot = dict(disp=w_disp.astype(np.float64), vel=w_vel.astype(np.float64),
               acc=np.zeros_like(ic64['disp']))
    _, pc, cc = O.nbody_adj(final, cot, ocosmo, o64)
    assert _cos(d.grad.cpu().numpy(), pc['disp']) >= 0.9999
    assert _cos(v.grad.cpu().numpy(), pc['vel']) >= 0.9999
    # CIC's weight gradient is piecewise constant (sign(-d), pm_util.py:144): one particle crossing a
    # cell face under float32 summation-order noise flips an O(1) term of a few elements, so an RMS
    # bound is a heavy-tailed statistic (the float32 oracle itself jumps from 6e-5 to 7e-3 under a 1e-6
    # cell perturbation of the ICs).  Bound the bulk instead: median and 99th percentile of |error|.
    for got, want in ((d.grad, pc['disp']), (v.grad, pc['vel'])):
        e = np.abs(got.cpu().numpy() - want) / _rms(want)
        # calibrated with the float32 oracle vs the float64 one over 8 draws of a 1e-6-cell IC
        # perturbation: median 1.1e-5 .. 9.7e-5, p99 2e-4 .. 9e-3 (rms 6e-5 .. 6e-3, max up to 0.6)
        assert np.median(e) <= 5e-4 and np.quantile(e, 0.99) <= 3e-2, (np.median(e), np.quantile(e, 0.99))
    # cosmology leaves: sums of float32 particle products over a chaotic 8-step 16^3 run;
    # compare with the float32 oracle's own distance from float64
    ic32 = dict(ic)
    final32 = O.nbody(ic32, ocosmo, oconf)
    cot32 = dict(disp=w_disp, vel=w_vel, acc=np.zeros_like(w_disp))
    _, _, cc32 = O.nbody_adj(final32, cot32, ocosmo, oconf)
    noise = abs(cc32['Omega_m'] / cc['Omega_m'] - 1)
    err = abs(Om.grad.item() / cc['Omega_m'] - 1)
    print('Omega_m cot rel err', err, 'float32-oracle rel err', noise)
    # (atomic and deterministic runs of ours differ from each other by ~1e-3 on this chaotic
    # 16^3 / 8-step configuration: that is the float32 noise level of this scalar; at 64^3 the float32 oracle
    # alone is 0.4 .. 2.2 % from the float64 one, profiles/r02_adjoint_noise_calibration.txt)
    assert err <= max(2e-2, 3 * noise)
    assert _cos(gt.grad.numpy(), cc['growth']) >= 0.9999


def test_nbody_adjoint_config1_size_vs_oracle():
    """BASELINE config 1 at its size (64^3 particles, 128^3 mesh, 10 leapfrog steps): the reverse-time
    adjoint (nbody.py:226-276) against the float64 oracle's nbody_adj.  Cosine >= 0.9999 for the disp and
    vel cotangents and for the growth-table cotangent; Omega_m cotangent within 5e-2 (a float32 sum over a
    chaotic run, bound calibrated with the float32 oracle; the bulk statistic of the particle cotangents as in
    the 16^3 test)."""
    pm, conf, oconf, cosmo, ocosmo, ic, ptcl = _ic(64, a_nbody_maxstep=0.1)
    assert conf.a_nbody_num == 10
    rng = np.random.default_rng(5)
    w_disp = rng.standard_normal(ic['disp'].shape).astype(np.float32)
    w_vel = rng.standard_normal(ic['disp'].shape).astype(np.float32)
    Om = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    gt = cosmo.growth.detach().clone().requires_grad_(True)
    c = cosmo.replace(Omega_m=Om, growth=gt)
    d = ptcl.disp.clone().requires_grad_(True)
    v = ptcl.vel.clone().requires_grad_(True)
    out, _ = pm.nbody(ptcl.replace(disp=d, vel=v), None, c, conf)
    obj = (out.disp * torch.from_numpy(w_disp).cuda()).sum() + (out.vel * torch.from_numpy(w_vel).cuda()).sum()
    obj.backward()

    o64 = O.Conf(1., (64,) * 3, mesh_shape=2, float_dtype=np.float64, a_nbody_maxstep=0.1)
    ic64 = dict(pmid=ic['pmid'], disp=ic['disp'].astype(np.float64), vel=ic['vel'].astype(np.float64))
    final = O.nbody(ic64, ocosmo, o64)
    cot = dict(disp=w_disp.astype(np.float64), vel=w_vel.astype(np.float64), acc=np.zeros_like(ic64['disp']))
    _, pc, cc = O.nbody_adj(final, cot, ocosmo, o64)
    for got, want in ((d.grad, pc['disp']), (v.grad, pc['vel'])):
        g = got.cpu().numpy()
        assert _cos(g, want) >= 0.9999
        e = np.abs(g - want) / _rms(want)
        print('cot err / rms: median', np.median(e), 'p99', np.quantile(e, 0.99), 'rms', _rms(e))
        assert np.median(e) <= 5e-4 and np.quantile(e, 0.99) <= 3e-2
    err = abs(Om.grad.item() / cc['Omega_m'] - 1)
    print('Omega_m cot rel err', err)
    # a float32 sum with cancellations over a chaotic run: the float32 ORACLE itself sits 2.2e-2, 4.2e-3, 4.8e-3 and
    # 1.9e-2 from the float64 one over four draws of a 1e-6-cell IC perturbation (tools/calib_adjoint_noise.py,
    # profiles/r02_adjoint_noise_calibration.txt); the tight checks are the cosines and the bulk statistics above
    assert err <= 5e-2
    assert _cos(gt.grad.numpy(), cc['growth']) >= 0.9999


def test_full_gradient_pipeline():
    """grads.py model: obj = var(scatter(nbody(lpt(linear_modes(white)))) - target).
    Gradient w.r.t. the real white-noise modes vs the oracle adjoint chain (cos >= 0.9999);
    gradient w.r.t. (A_s, n_s, Omega_m, Omega_b, h) vs float64 finite differences of the
    oracle pipeline (cos >= 0.9999)."""
    pm = _pm()
    n = 16
    kw = dict(a_start=1 / 16, a_nbody_maxstep=1 / 8)
    conf, oconf = _confs((n, n, n), **kw)
    o64 = O.Conf(1., (n,) * 3, mesh_shape=2, float_dtype=np.float64, **kw)
    names = ('A_s_1e9', 'n_s', 'Omega_m', 'Omega_b', 'h')
    vals = dict(A_s_1e9=2.0, n_s=0.96, Omega_m=0.3, Omega_b=0.05, h=0.7)
    white = O.white_noise(1, oconf, real=True)
    tgt_white = O.white_noise(0, oconf, real=True)

    def omodel(w, p, conf_):
        c = O.boltzmann(O.Cosmo(conf_, **p), conf_)
        ic = O.lpt(O.linear_modes(w.astype(conf_.float_dtype), c, conf_), c, conf_)
        out = O.nbody(ic, c, conf_)
        return O.scatter(out['pmid'], out['disp'], conf_), out, c

    target, _, _ = omodel(tgt_white, vals, o64)

    # ---- ours
    theta = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in vals.items()}
    cosmo = pm.boltzmann(pm.Cosmology(conf, **theta), conf)
    w = torch.from_numpy(white).cuda().requires_grad_(True)
    modes = pm.linear_modes(w, cosmo, conf)
    ptcl, obs = pm.lpt(modes, cosmo, conf)
    ptcl, obs = pm.nbody(ptcl, obs, cosmo, conf)
    dens = pm.scatter(ptcl, conf)
    obj = (dens - torch.from_numpy(target.astype(np.float32)).cuda()).var(unbiased=False)
    obj.backward()

    # ---- oracle: objective value
    dens64, out64, c64 = omodel(white, vals, o64)
    obj64 = np.var(dens64 - target)
    np.testing.assert_allclose(obj.item(), obj64, rtol=1e-3)

    # ---- oracle: white-noise gradient by the adjoint chain in float64
    dens_cot = 2 * (dens64 - target - np.mean(dens64 - target)) / dens64.size
    dcot, _ = O.scatter_adj(out64['pmid'], out64['disp'], o64, dens_cot)
    cot = dict(disp=dcot, vel=np.zeros_like(dcot), acc=np.zeros_like(dcot))
    _, pc, _ = O.nbody_adj(out64, cot, c64, o64)
    gw = O.lpt_vjp_modes(white.astype(np.float64), c64, o64, pc['disp'], pc['vel'])
    assert _cos(w.grad.cpu().numpy(), gw) >= 0.9999
    assert _rms(w.grad.cpu().numpy() - gw) <= 2e-2 * _rms(gw)

    # ---- oracle: cosmology gradient by central finite differences of the whole model
    fd = []
    for nme in names:
        h = 1e-6 * abs(vals[nme])   # small: CIC is only piecewise smooth
        pp = dict(vals); pp[nme] += h
        pm_ = dict(vals); pm_[nme] -= h
        fp = np.var(omodel(white, pp, o64)[0] - target)
        fm = np.var(omodel(white, pm_, o64)[0] - target)
        fd.append((fp - fm) / (2 * h))
    got = [theta[nme].grad.item() for nme in names]
    assert _cos(got, fd) >= 0.9999, (got, fd)
    np.testing.assert_allclose(got, fd, rtol=2e-2)


# the last three shapes have more tiles than resident CTAs (persistent loop + cp.async prefetch)
@pytest.mark.parametrize('shape', [(64, 8, 10), (128, 6, 9), (256, 4, 6), (512, 4, 4), (1024, 2, 4), (2048, 2, 4),
                                   (256, 80, 130), (512, 40, 130), (1024, 24, 130), (2048, 24, 130)],
                         ids=lambda shp: 'x'.join(str(n) for n in shp))
def test_fused_xpass_vs_cufft3d(shape):
    """csrc/xpass.cu / xpass16.cu (on-chip FFT along x fused with the k-space algebra) against the
    plain pipeline cuFFT-3D -> pmwd_kspace_force(_adj) -> cuFFT-3D, forward and adjoint, incl. the
    y-slab form used after the distributed all-to-all.  Tolerance: 2e-6 of the field's rms."""
    import ctypes as C
    from pmwd_b200 import _lib
    lib = _lib.lib()
    nx, ny, nz = shape
    cell, scale = 0.7, 0.37
    g = torch.Generator(device='cuda').manual_seed(0)
    rho = torch.randn(shape, device='cuda', generator=g)
    st = _lib.stream_ptr()
    shp = _lib.shape_arr(shape)
    # reference: full 3-D transforms + the standalone fused k-space kernel
    spec = torch.fft.rfftn(rho).contiguous()
    ref_g = [torch.empty_like(spec) for _ in range(3)]
    arr = (C.c_void_p * 3)(*[t.data_ptr() for t in ref_g])
    _lib.check(lib.pmwd_kspace_force(st, 3, shp, cell, scale, _lib.ptr(spec), arr), 'ks')
    ref_F = [torch.fft.irfftn(t, s=shape, norm='forward') for t in ref_g]
    # fused: 2-D transforms over (y, z) + x-pass
    s2 = torch.fft.rfft2(rho).contiguous()
    out = [torch.empty_like(s2) for _ in range(3)]
    arr = (C.c_void_p * 3)(*[t.data_ptr() for t in out])
    _lib.check(lib.pmwd_xpass_force(st, shp, 0, ny, cell, scale, _lib.ptr(s2), arr), 'xpass')
    for a in range(3):
        F = torch.fft.irfft2(out[a], s=(ny, nz), norm='forward')
        err = _rms((F - ref_F[a]).cpu().numpy())
        assert err <= 2e-6 * _rms(ref_F[a].cpu().numpy()) + 1e-12, (a, err, _rms(ref_F[a].cpu().numpy()))
    # y-slab form: rows y0 .. y0+nyl-1 only
    y0, nyl = ny // 2, ny - ny // 2
    s2s = s2[:, y0:y0 + nyl].contiguous()
    outs = [torch.empty_like(s2s) for _ in range(3)]
    arr = (C.c_void_p * 3)(*[t.data_ptr() for t in outs])
    _lib.check(lib.pmwd_xpass_force(st, shp, y0, nyl, cell, scale, _lib.ptr(s2s), arr), 'xpass slab')
    nzc = nz // 2 + 1
    import os
    reg = (256, 512, 1024) if os.environ.get('PMWD_XPASS16_2048') == '0' else (256, 512, 1024, 2048)
    same_kernel = nx not in reg or (nyl * nzc) % 2 == (ny * nzc) % 2
    for a in range(3):
        if same_kernel:     # per-column arithmetic does not depend on the slab
            assert torch.equal(outs[a], out[a][:, y0:y0 + nyl])
        else:               # odd column count: the slab falls back to the radix-4 kernel
            d = (outs[a] - out[a][:, y0:y0 + nyl]).abs().max().item()
            assert d <= 1e-5 * out[a].abs().max().item()
    # adjoint: three inputs -> one output
    V = [torch.randn(shape, device='cuda', generator=g) for _ in range(3)]
    Vk = [torch.fft.rfftn(v).contiguous() for v in V]
    ref_o = torch.empty_like(spec)
    arr = (C.c_void_p * 3)(*[t.data_ptr() for t in Vk])
    _lib.check(lib.pmwd_kspace_force_adj(st, 3, shp, cell, scale, arr, _lib.ptr(ref_o)), 'ks adj')
    ref_r = torch.fft.irfftn(ref_o, s=shape, norm='forward')
    V2 = [torch.fft.rfft2(v).contiguous() for v in V]
    o2 = torch.empty_like(s2)
    arr = (C.c_void_p * 3)(*[t.data_ptr() for t in V2])
    _lib.check(lib.pmwd_xpass_force_adj(st, shp, 0, ny, cell, scale, arr, _lib.ptr(o2)), 'xpass adj')
    r = torch.fft.irfft2(o2, s=(ny, nz), norm='forward')
    assert _rms((r - ref_r).cpu().numpy()) <= 2e-6 * _rms(ref_r.cpu().numpy())


def test_fused_xpass_2048_radix4_fallback():
    """nx = 2048 defaults to the register x-pass on two-CTA clusters (csrc/xpass16.cu); the radix-4
    shared-memory kernel (csrc/xpass.cu, PMWD_XPASS16_2048=0) stays as its A/B partner and runs through
    the same parity test in a process of its own (the switch is read once)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PMWD_XPASS16_2048='0')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_gravity.py'), '-q', '-x',
                        '-m', 'gpu', '-k', 'test_fused_xpass_vs_cufft3d and 2048x'], env=env, cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert '2 passed' in r.stdout


@pytest.mark.parametrize('splits', [((0, 0.5), (1, 0.5)), ((0.5, 0), (0.5, 1))])
def test_nbody_step_host_matches_nbody_step(splits):
    """Host-array entry point == copy up, nbody_step, copy down: displacements bit-identical, velocities and
    accelerations up to the scatter's summation order; also in place (out = host) over several steps and for
    a non-default splitting (plain route inside)."""
    pm, conf, oconf, cosmo, ocosmo, ic, ptcl = _ic(32, a_nbody_maxstep=1 / 16, symp_splits=splits)
    a = conf.a_nbody
    p0, _ = pm.nbody_init(a[0], ptcl, None, cosmo, conf)
    host = {k: getattr(p0, k).cpu().pin_memory() for k in ('pmid', 'disp', 'vel', 'acc')}
    ref = p0
    for i in range(3):
        ref, _ = pm.nbody_step(a[i], a[i + 1], ref, None, cosmo, conf)
        host = pm.nbody_step_host(a[i], a[i + 1], host, cosmo, conf, out=host if i else None)
        torch.cuda.synchronize()
        assert torch.equal(host['pmid'], ref.pmid.cpu())
        if i == 0:
            assert torch.equal(host['disp'], ref.disp.cpu())
        for k in ('disp', 'vel', 'acc'):
            r = getattr(ref, k).cpu()
            assert (host[k] - r).abs().max().item() <= 1e-5 * r.abs().max().item(), (i, k)
    with pytest.raises(ValueError):
        pm.nbody_step_host(a[0], a[1], dict(host, disp=host['disp'].cuda()), cosmo, conf)


def test_nbody_step_host_acc_mirror():
    """The device keeps the accelerations it computed: a host acc array that is the one the previous call
    wrote is not uploaded again, one that was modified through torch (version counter) or replaced is,
    and acc_resident=False always uploads.  All routes give the result of a call with explicit uploads."""
    import importlib
    nb = importlib.import_module('pmwd_b200.nbody')      # (pmwd_b200.nbody is also the function)
    pm, conf, oconf, cosmo, ocosmo, ic, ptcl = _ic(32, a_nbody_maxstep=1 / 16)
    a = conf.a_nbody
    p0, _ = pm.nbody_init(a[0], ptcl, None, cosmo, conf)
    first = {k: getattr(p0, k).cpu().pin_memory() for k in ('pmid', 'disp', 'vel', 'acc')}
    nb.nbody_host_release()
    s1 = pm.nbody_step_host(a[0], a[1], first, cosmo, conf)
    torch.cuda.synchronize()
    state = {k: s1[k].clone().pin_memory() for k in s1}

    def step2(src, **kw):
        o = pm.nbody_step_host(a[1], a[2], src, cosmo, conf, **kw)
        torch.cuda.synchronize()
        return {k: o[k].clone() for k in ('disp', 'vel', 'acc')}

    def same(x, y):       # two runs of step 1 differ by the atomic deposit's summation order
        for k in ('disp', 'vel', 'acc'):
            assert (x[k] - y[k]).abs().max().item() <= 1e-5 * y[k].abs().max().item(), k

    want = step2(state, acc_resident=False)              # explicit upload of everything
    # (a) the array the previous call wrote: acc comes from the device mirror
    nb.nbody_host_release()
    s1 = pm.nbody_step_host(a[0], a[1], first, cosmo, conf)
    m = nb._host_mirrors[str(torch.device('cuda', torch.cuda.current_device()))]
    assert m.acc_tag == (s1['acc'].data_ptr(), tuple(s1['acc'].shape), s1['acc']._version)
    same(step2(s1), want)
    # (b) acc modified on the host through torch: the tag no longer matches, the modified array goes up
    nb.nbody_host_release()
    s1 = pm.nbody_step_host(a[0], a[1], first, cosmo, conf)
    torch.cuda.synchronize()
    s1['acc'].mul_(2.0)
    mod = dict(state, acc=(state['acc'] * 2.0).pin_memory())
    same(step2(s1), step2(mod, acc_resident=False))
    # (c) a stale mirror must not leak into a call with another acc array
    same(step2(state), want)
