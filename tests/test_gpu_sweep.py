"""GPU parity of the tiled ("tile sweep") CIC deposit (csrc/scatter_sweep.cu) against the oracle's
scatter (pmwd/scatter.py:60-83) and against the per-particle RED kernel: freshly sorted storage,
stale sorts (stragglers), 3 channels, anisotropic meshes with odd tile shapes.
Tolerance: per-cell density rel. err <= 1e-5 (of max(density, 1)), the north-star figure."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


def _setup(shape, sigma, seed=0, spacing=1., **kw):
    import pmwd_b200 as pm
    conf = pm.Configuration(spacing, shape, mesh_shape=2, **kw)
    oconf = O.Conf(spacing, shape, mesh_shape=2)
    pmid, disp, _, _ = O.gen_grid(oconf)
    rng = np.random.default_rng(seed)
    disp = (disp + sigma * conf.cell_size * rng.standard_normal(disp.shape)).astype(np.float32)
    ptcl = pm.Particles(conf, torch.from_numpy(np.ascontiguousarray(pmid)).cuda(), torch.from_numpy(disp).cuda(),
                        vel=torch.zeros(disp.shape, device='cuda'))
    return pm, conf, oconf, pmid, disp, ptcl


def _sweep_scatter(store, conf, val=None, nch=1):
    from pmwd_b200 import _lib
    from pmwd_b200.gravity import _force_desc
    a = store.arrays
    desc = _force_desc(a['pmid'], conf)
    dev = a['disp'].device
    # garbage-filled outputs: the sweep must overwrite every cell
    m = [torch.full(tuple(conf.mesh_shape), 7.5, device=dev) for _ in range(nch)]
    arg = store.sweep_arg()
    assert arg is not None
    _lib.check(_lib.lib().pmwd_scatter_sweep(
        _lib.stream_ptr(dev), C.byref(desc), arg, _lib.ptr(a['pmid']), _lib.ptr(a['disp']),
        _lib.ptr(val), float(conf.mesh_size / conf.ptcl_num), nch, _lib.ptr(m[0]),
        _lib.ptr(m[1]) if nch == 3 else None, _lib.ptr(m[2]) if nch == 3 else None), 'pmwd_scatter_sweep')
    torch.cuda.synchronize()
    return m


def _close(got, want):
    got = got.cpu().numpy().astype(np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert abs(got.sum() - want.sum()) <= 1e-6 * abs(want).sum() + 1e-3
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert rel.max() <= 1e-5, rel.max()


@pytest.mark.parametrize('shape, sigma', [((32, 32, 32), 0.3), ((32, 32, 32), 3.0), ((12, 10, 18), 0.4),
                                          ((8, 8, 512), 0.5), ((40, 6, 34), 0.2), ((16, 16, 16), 40.0),
                                          ((64, 16, 48), 1.0)],
                         ids=lambda v: 'x'.join(str(n) for n in v) if isinstance(v, tuple) else str(v))
def test_sweep_scatter_fresh_sort_vs_oracle(shape, sigma):
    """The store sorts its arrays into the tile order when it is set up: every particle goes through
    the shared-memory tiles (no stragglers), whatever the displacement amplitude and tile shape."""
    from pmwd_b200.nbody import _store_from
    pm, conf, oconf, pmid, disp, ptcl = _setup(shape, sigma)
    store = _store_from(ptcl, conf)
    assert store.sweep is not None and store.sweep.ok and store.reorders == 1
    st = store.sweep.status.cpu()
    assert int(st[0]) & 0xffffffff == 0 and int(st[1]) == conf.ptcl_num      # the table is an exact partition
    dens, = _sweep_scatter(store, conf)
    _close(dens, O.scatter(pmid, disp, oconf))
    assert store.sweep.stragglers() == 0


@pytest.mark.parametrize('spacing, force_div', [(0.7, False), (1.0, True), (0.1, False)])
def test_sweep_scatter_cell_size_paths(spacing, force_div, monkeypatch):
    """disp / cell (pm_util.py:133): cells that are not a power of two take the IEEE division, powers of two
    the exact reciprocal; both must give the oracle's deposit, also with partly filled warps (fresh and stale)."""
    from pmwd_b200.nbody import _store_from
    if force_div:
        monkeypatch.setenv('PMWD_SWEEP_DIV', '1')
    pm, conf, oconf, pmid, disp, ptcl = _setup((32, 16, 32), 1.5, spacing=spacing)
    store = _store_from(ptcl, conf)
    assert store.sweep is not None and store.sweep.ok
    dens, = _sweep_scatter(store, conf)
    _close(dens, O.scatter(pmid, disp, oconf))
    assert store.sweep.stragglers() == 0
    g = torch.Generator(device='cuda').manual_seed(5)
    store.arrays['disp'] += 0.5 * conf.cell_size * torch.randn(store.arrays['disp'].shape, device='cuda', generator=g)
    dens, = _sweep_scatter(store, conf)
    _close(dens, O.scatter(pmid, store.lagrangian('disp').cpu().numpy(), oconf))


@pytest.mark.parametrize('shape, sigma', [((32, 32, 32), 2.0), ((24, 16, 64), 5.0), ((8, 8, 512), 6.0)],
                         ids=lambda v: 'x'.join(str(n) for n in v) if isinstance(v, tuple) else str(v))
def test_sweep_scatter_sorted_and_stale_order(shape, sigma):
    """After a re-sort (table from the sort's keys) no particle is a straggler; after the particles
    have moved on without a re-sort the result is still exact and only the straggler count grows."""
    from pmwd_b200.nbody import _store_from
    pm, conf, oconf, pmid, disp, ptcl = _setup(shape, sigma)
    store = _store_from(ptcl, conf)
    assert store.sweep.ok
    dens, = _sweep_scatter(store, conf)
    _close(dens, O.scatter(pmid, disp, oconf))
    assert store.sweep.stragglers() == 0
    # drift by up to ~1.5 cells without re-sorting
    g = torch.Generator(device='cuda').manual_seed(3)
    kick = 0.6 * conf.cell_size * torch.randn(store.arrays['disp'].shape, device='cuda', generator=g)
    store.arrays['disp'] += kick
    dens, = _sweep_scatter(store, conf)
    lag = store.lagrangian('disp').cpu().numpy()
    _close(dens, O.scatter(pmid, lag, oconf))
    n_strag = store.sweep.stragglers()
    print('stragglers after drift', n_strag, 'of', conf.ptcl_num)
    assert 0 < n_strag < conf.ptcl_num


def test_sweep_scatter_three_channels_vs_red_kernel():
    """The adjoint's V_i = scatter(pi_i) (gather.py:113): three sweeps against pmwd_scatter_soa."""
    from pmwd_b200 import _lib
    from pmwd_b200.gravity import _force_desc
    from pmwd_b200.nbody import _store_from
    pm, conf, oconf, pmid, disp, ptcl = _setup((32, 32, 32), 2.0)
    store = _store_from(ptcl, conf)
    a = store.arrays
    g = torch.Generator(device='cuda').manual_seed(1)
    pi = torch.randn(a['disp'].shape, device='cuda', generator=g)
    dens, = _sweep_scatter(store, conf)          # records the stragglers (none) like force_adj does
    V = _sweep_scatter(store, conf, val=pi, nch=3)
    desc = _force_desc(a['pmid'], conf)
    ref = [torch.zeros(tuple(conf.mesh_shape), device='cuda') for _ in range(3)]
    _lib.check(_lib.lib().pmwd_scatter_soa(_lib.stream_ptr(), C.byref(desc), _lib.ptr(a['pmid']), _lib.ptr(a['disp']),
                                           _lib.ptr(pi), 0.0, 3, _lib.ptr(ref[0]), _lib.ptr(ref[1]), _lib.ptr(ref[2])),
               'pmwd_scatter_soa')
    for c in range(3):
        scale = ref[c].abs().max().item()
        assert (V[c] - ref[c]).abs().max().item() <= 2e-6 * scale
    # and against the oracle's gather VJP mesh cotangent for one channel
    lag_disp, lag_pi = store.lagrangian('disp').cpu().numpy(), None
    pi_lag = torch.empty_like(pi)
    pi_lag[store.lag.long()] = pi
    _, mc = O.gather_adj(pmid, lag_disp, oconf, np.zeros(oconf.mesh_shape, np.float32), pi_lag[:, 1].cpu().numpy())
    assert np.abs(V[1].cpu().numpy() - mc).max() <= 2e-6 * np.abs(mc).max()


def test_sweep_handles_arbitrary_input_order():
    """The caller's particle order does not matter (the store sorts its private copies): a randomly
    permuted input gives the same density and the outputs come back in the caller's order."""
    from pmwd_b200.nbody import _store_from
    pm, conf, oconf, pmid, disp, ptcl = _setup((16, 16, 16), 0.8)
    perm = torch.randperm(conf.ptcl_num, device='cuda', generator=torch.Generator(device='cuda').manual_seed(0))
    shuffled = pm.Particles(conf, ptcl.pmid[perm].contiguous(), ptcl.disp[perm].contiguous(), vel=ptcl.vel)
    store = _store_from(shuffled, conf)
    assert store.sweep is not None and store.sweep.ok and store.sweep_arg() is not None
    dens, = _sweep_scatter(store, conf)
    _close(dens, O.scatter(pmid, disp, oconf))
    assert torch.equal(store.lagrangian('disp'), shuffled.disp)


@pytest.mark.parametrize('tiled', [True, False])
def test_nbody_same_result_with_and_without_tiles(tiled):
    """The integrator with the tiled deposit (default) and with the RED kernel (scatter_tiled=False)
    give the same trajectory up to float32 summation order: both within 1e-4 cell of the oracle."""
    import pmwd_b200 as pm
    n = 32
    kw = dict(a_nbody_maxstep=1 / 8)
    conf = pm.Configuration(1., (n,) * 3, mesh_shape=2, scatter_tiled=tiled, reorder_min_disp=0.5, **kw)
    oconf = O.Conf(1., (n,) * 3, mesh_shape=2, **kw)
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    ocosmo = O.boltzmann(O.SimpleLCDM(oconf), oconf).replace(growth=cosmo.growth.numpy())
    ic = O.lpt(O.linear_modes(O.white_noise(0, oconf), ocosmo, oconf), ocosmo, oconf)
    ptcl = pm.Particles(conf, torch.from_numpy(ic['pmid']).cuda(), torch.from_numpy(ic['disp']).cuda(),
                        vel=torch.from_numpy(ic['vel']).cuda())
    out, _ = pm.nbody(ptcl, None, cosmo, conf)
    ref = O.nbody(dict(ic), ocosmo, oconf)
    err = np.abs(out.disp.cpu().numpy() - ref['disp']) / conf.cell_size
    assert np.sqrt(np.mean(err ** 2)) <= 1e-4 and np.quantile(err, 0.999) <= 1e-4


@pytest.mark.parametrize('x0, mx, h', [(16, 32, 8), (48, 16, 8), (0, 32, 4)])
def test_sweep_scatter_on_a_slab_descriptor(x0, mx, h):
    """Multi-GPU composition on one GPU: the mesh array is an x-slab of `mx` planes + `h` halo planes on either
    side (possibly wrapping around the box), particles whose stencil leaves it are dropped plane by plane
    (enmesh's rule for a smaller target mesh, pm_util.py:147-154).  With the storage sorted for THAT descriptor
    the tiled deposit must equal the per-particle RED kernel on the same descriptor -- freshly sorted, after a
    drift, and for three channels -- and must refuse a descriptor with another halo width."""
    from pmwd_b200 import _lib
    from pmwd_b200.nbody import _Store
    from pmwd_b200.scatter import make_desc
    pm, conf, oconf, pmid, disp, ptcl = _setup((32, 16, 32), 1.0)
    Mx, My, Mz = conf.mesh_shape
    cell32 = float(np.float32(conf.cell_size))

    def desc_fn(pmid_, hh=h):
        return make_desc(conf, pmid_, (mx + 2 * hh, My, Mz), 1, ((x0 - hh) * cell32, 0.0, 0.0), None)

    store = _Store(conf, dict(pmid=ptcl.pmid.clone(), disp=ptcl.disp.clone(), vel=torch.zeros_like(ptcl.disp),
                              acc=torch.zeros_like(ptcl.disp)))
    store.desc_fn = desc_fn
    store.slab_sweep = True
    store.reorder()
    assert store.sweep is not None and store.sweep.ok
    lib, st = _lib.lib(), _lib.stream_ptr()
    g = torch.Generator(device='cuda').manual_seed(2)

    def both(nch):
        a = store.arrays
        desc = desc_fn(a['pmid'])
        assert store.sweep.usable(desc)
        assert not store.sweep.usable(desc_fn(a['pmid'], h + 4))
        shape = (mx + 2 * h, My, Mz)
        val = torch.randn(a['disp'].shape, device='cuda', generator=g) if nch == 3 else None
        got = [torch.full(shape, -3.25, device='cuda') for _ in range(nch)]
        ref = [torch.zeros(shape, device='cuda') for _ in range(nch)]
        extra = lambda m: [_lib.ptr(m[1]) if nch == 3 else None, _lib.ptr(m[2]) if nch == 3 else None]
        _lib.check(lib.pmwd_scatter_sweep(st, C.byref(desc), store.sweep.arg(), _lib.ptr(a['pmid']), _lib.ptr(a['disp']),
                                          _lib.ptr(val), 0.125, nch, _lib.ptr(got[0]), *extra(got)), 'pmwd_scatter_sweep')
        _lib.check(lib.pmwd_scatter_soa(st, C.byref(desc), _lib.ptr(a['pmid']), _lib.ptr(a['disp']), _lib.ptr(val), 0.125,
                                        nch, _lib.ptr(ref[0]), *extra(ref)), 'pmwd_scatter_soa')
        torch.cuda.synchronize()
        for c in range(nch):
            scale = max(ref[c].abs().max().item(), 1.0)
            assert (got[c] - ref[c]).abs().max().item() <= 2e-6 * scale, (nch, c)
        assert ref[0].abs().sum().item() > 0

    both(1)
    store.arrays['disp'] += 0.6 * conf.cell_size * torch.randn(store.arrays['disp'].shape, device='cuda', generator=g)
    both(1)
    both(3)


@pytest.mark.parametrize('shape, sigma', [((32, 32, 32), 2.0), ((8, 8, 512), 1.0), ((12, 10, 18), 0.7), ((64, 16, 48), 3.0)],
                         ids=lambda v: 'x'.join(str(n) for n in v) if isinstance(v, tuple) else str(v))
def test_sweep_det_is_bitwise_reproducible_and_matches_oracle(shape, sigma):
    """The deterministic variant of the tiled deposit (one warp per (y, z) tile over all planes, per-tile halo
    arrays added in a fixed order, no float atomics): two runs agree bit for bit although the tiles are handed to
    the warps dynamically, the density matches the oracle (scatter.py:60-83) to <= 1e-5, three channels match the
    RED kernel, nothing goes through the straggler path for freshly sorted storage -- and a stale order is still
    deposited correctly but reported."""
    from pmwd_b200 import _lib
    from pmwd_b200.gravity import _force_desc
    from pmwd_b200.nbody import _store_from
    pm, conf, oconf, pmid, disp, ptcl = _setup(shape, sigma, scatter_mode='deterministic')
    store = _store_from(ptcl, conf)
    assert store.sweep is not None and store.sweep.ok and store.sweep.det and store.det_sweep
    lib, st = _lib.lib(), _lib.stream_ptr()

    def run(nch, val=None):
        a = store.arrays
        desc = _force_desc(a['pmid'], conf)
        m = [torch.full(tuple(conf.mesh_shape), 7.5, device='cuda') for _ in range(nch)]
        _lib.check(lib.pmwd_scatter_sweep_det(
            st, C.byref(desc), store.sweep.arg(), _lib.ptr(a['pmid']), _lib.ptr(a['disp']), _lib.ptr(val),
            float(conf.mesh_size / conf.ptcl_num), nch, _lib.ptr(m[0]),
            _lib.ptr(m[1]) if nch == 3 else None, _lib.ptr(m[2]) if nch == 3 else None), 'pmwd_scatter_sweep_det')
        torch.cuda.synchronize()
        return m

    d1, = run(1)
    d2, = run(1)
    assert torch.equal(d1, d2)
    _close(d1, O.scatter(pmid, disp, oconf))
    g = torch.Generator(device='cuda').manual_seed(4)
    pi = torch.randn(store.arrays['disp'].shape, device='cuda', generator=g)
    V1, V2 = run(3, pi), run(3, pi)
    a = store.arrays
    desc = _force_desc(a['pmid'], conf)
    ref = [torch.zeros(tuple(conf.mesh_shape), device='cuda') for _ in range(3)]
    _lib.check(lib.pmwd_scatter_soa(st, C.byref(desc), _lib.ptr(a['pmid']), _lib.ptr(a['disp']), _lib.ptr(pi), 0.0, 3,
                                    _lib.ptr(ref[0]), _lib.ptr(ref[1]), _lib.ptr(ref[2])), 'pmwd_scatter_soa')
    torch.cuda.synchronize()
    for c in range(3):
        assert torch.equal(V1[c], V2[c])
        assert (V1[c] - ref[c]).abs().max().item() <= 2e-6 * max(ref[c].abs().max().item(), 1.0)
    assert store.sweep.det_violations() == 0
    # stale order: still the right density, but flagged
    store.arrays['disp'] += 0.7 * conf.cell_size * torch.randn(store.arrays['disp'].shape, device='cuda', generator=g)
    d3, = run(1)
    _close(d3, O.scatter(pmid, store.lagrangian('disp').cpu().numpy(), oconf))
    assert store.sweep.det_violations() > 0


def test_deterministic_nbody_through_the_tiled_deposit_is_reproducible():
    """Deterministic mode inside the integrator: re-sort before every force, reproducible tiled deposit.  Two runs
    agree bit for bit (forward and adjoint), and with the cell-sorted scatter of scatter_det.cu (PMWD_SWEEP_DET=0)
    to float32 summation-order accuracy."""
    import os
    import pmwd_b200 as pm
    oconf = O.Conf(1., (32, 32, 32), mesh_shape=2, a_nbody_maxstep=1 / 8)
    conf = pm.Configuration(1., (32, 32, 32), mesh_shape=2, a_nbody_maxstep=1 / 8, scatter_mode='deterministic')
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    ic, _ = pm.lpt(pm.linear_modes(pm.white_noise(0, conf, real=True), cosmo, conf), cosmo, conf)
    outs = []
    for k in range(3):
        if k == 2:
            os.environ['PMWD_SWEEP_DET'] = '0'
        try:
            out, _ = pm.nbody(ic, None, cosmo, conf)
            g = torch.Generator(device='cuda').manual_seed(9)
            cot = pm.Particles(conf, out.pmid, torch.randn(out.disp.shape, device='cuda', generator=g),
                               vel=torch.zeros_like(out.disp))
            _, pc, _ = pm.nbody_adj(out, cot, None, cosmo, conf)
        finally:
            os.environ.pop('PMWD_SWEEP_DET', None)
        torch.cuda.synchronize()
        outs.append((out.disp.clone(), out.vel.clone(), pc.disp.clone(), pc.vel.clone()))
    for x, y in zip(outs[0], outs[1]):
        assert torch.equal(x, y)
    for x, y in zip(outs[0][:2], outs[2][:2]):
        assert (x - y).abs().max().item() <= 1e-3 * y.abs().max().item()
    a_, b_ = outs[0][2].double().flatten(), outs[2][2].double().flatten()
    assert float(a_ @ b_ / torch.sqrt((a_ @ a_) * (b_ @ b_))) >= 0.999
