"""GPU tests of the fused passes added in round 2, each against the unfused sequence it replaces
(bit-exact: same float32 operations in the same order) or against torch autograd of the plain expression:

* pmwd_kick_kick_drift_adj == pmwd_kick_drift_adj(kick) followed by pmwd_kick_drift_adj(kick + drift)
  (pmwd/nbody.py:49-99, 143-162);
* pmwd_force_adj's acc (evaluated inside the weight-gradient gather) == pmwd_force's acc (gravity.py:47-72);
* pmwd_lpt_source2 / pmwd_lpt_displace and their VJP kernels (pmwd/lpt.py:40-76, 203-208);
* pmwd_slab_owner (owner rank + halo width) == the tensor formula of pmwd_b200/migrate.py / dist.py."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(shape, device='cuda', generator=g)


def test_kick_kick_drift_adj_equals_two_calls():
    from pmwd_b200 import _lib
    lib = _lib.lib()
    n = 3 * 100003
    base = {k: _rand(n, i) for i, k in enumerate(('disp', 'vel', 'acc', 'xi', 'pi', 'alpha'))}
    K0, K, D = 0.0123, -0.0345, 0.0567
    st = _lib.stream_ptr()

    a = {k: v.clone() for k, v in base.items()}
    s0 = torch.zeros(2, dtype=torch.float64, device='cuda')
    s1 = torch.zeros(2, dtype=torch.float64, device='cuda')
    _lib.check(lib.pmwd_kick_drift_adj(st, n, _lib.ptr(a['disp']), _lib.ptr(a['vel']), _lib.ptr(a['acc']), _lib.ptr(a['xi']),
                                       _lib.ptr(a['pi']), _lib.ptr(a['alpha']), K0, 0.0, 1, 0, _lib.ptr(s0)), 'kd0')
    _lib.check(lib.pmwd_kick_drift_adj(st, n, _lib.ptr(a['disp']), _lib.ptr(a['vel']), _lib.ptr(a['acc']), _lib.ptr(a['xi']),
                                       _lib.ptr(a['pi']), _lib.ptr(a['alpha']), K, D, 1, 1, _lib.ptr(s1)), 'kd1')
    b = {k: v.clone() for k, v in base.items()}
    t0 = torch.zeros(2, dtype=torch.float64, device='cuda')
    t1 = torch.zeros(2, dtype=torch.float64, device='cuda')
    _lib.check(lib.pmwd_kick_kick_drift_adj(st, n, _lib.ptr(b['disp']), _lib.ptr(b['vel']), _lib.ptr(b['acc']),
                                            _lib.ptr(b['xi']), _lib.ptr(b['pi']), _lib.ptr(b['alpha']), K0, K, D,
                                            _lib.ptr(t0), _lib.ptr(t1)), 'kkd')
    torch.cuda.synchronize()
    for k in base:
        assert torch.equal(a[k], b[k]), k
    np.testing.assert_allclose(t0[0].item(), s0[0].item(), rtol=1e-12)
    np.testing.assert_allclose(t1.cpu().numpy(), s1.cpu().numpy(), rtol=1e-12)
    # and against the plain expressions (nbody.py:87-97, 56-64) in float64 for the sums
    ref_pa = (base['pi'].double() * base['acc'].double()).sum().item()
    np.testing.assert_allclose(t0[0].item(), ref_pa, rtol=1e-10)


def test_force_adj_acc_equals_force_acc():
    import pmwd_b200 as pm
    from pmwd_b200.gravity import force_into, force_adj_into
    n = 32
    conf = pm.Configuration(1., (n, n, n), mesh_shape=2, scatter_mode='deterministic')
    ptcl = pm.Particles.gen_grid(conf)
    disp = (ptcl.disp + 1.7 * _rand(ptcl.disp.shape, 3)).contiguous()
    pi = _rand(ptcl.disp.shape, 4)
    acc0, acc1, alpha = torch.empty_like(disp), torch.empty_like(disp), torch.empty_like(disp)
    force_into(ptcl.pmid, disp, 0.3, conf, acc0)
    force_adj_into(ptcl.pmid, disp, 0.3, conf, pi, acc1, alpha)
    torch.cuda.synchronize()
    assert torch.equal(acc0, acc1)          # deterministic deposit -> identical meshes -> identical gathers


def test_lpt_fused_kernels_vs_torch():
    from pmwd_b200.lpt import _Source2, _Displace
    n = 50021
    s = [_rand(n, 10 + i).requires_grad_(True) for i in range(6)]
    L = _Source2.apply(*s)
    a, b, c, d, e, f = [t.detach().clone().requires_grad_(True) for t in s]
    ref = (((a * c + a * b) + b * c) - d * d - e * e) - f * f
    assert torch.equal(L.detach(), ref.detach())                 # same order, --fmad=false
    w = _rand(n, 20)
    L.backward(w)
    ref.backward(w)
    for got, want in zip(s, (a, b, c, d, e, f)):
        assert torch.allclose(got.grad, want.grad, rtol=1e-6, atol=1e-6)

    disp0, vel0 = _rand((n, 3), 30), _rand((n, 3), 31)
    g = [_rand(n, 40 + i).requires_grad_(True) for i in range(6)]
    fac = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (0.7, -1.3, 0.21, 0.42)]
    d, v = _Displace.apply(disp0, vel0, *fac, *g)
    g2 = [t.detach().clone().requires_grad_(True) for t in g]
    fac2 = [t.detach().clone().requires_grad_(True) for t in fac]
    f32 = [t.to(torch.float32).cuda() for t in fac2]
    dref = torch.stack([(disp0[:, i] + f32[0] * g2[i]) + f32[2] * g2[3 + i] for i in range(3)], dim=-1)
    vref = torch.stack([(vel0[:, i] + f32[1] * g2[i]) + f32[3] * g2[3 + i] for i in range(3)], dim=-1)
    assert torch.equal(d.detach(), dref.detach()) and torch.equal(v.detach(), vref.detach())
    wd, wv = _rand((n, 3), 50), _rand((n, 3), 51)
    ((d * wd).sum() + (v * wv).sum()).backward()
    ((dref * wd).sum() + (vref * wv).sum()).backward()
    for got, want in zip(g, g2):
        assert torch.allclose(got.grad, want.grad, rtol=1e-6, atol=1e-6)
    for got, want in zip(fac, fac2):
        np.testing.assert_allclose(got.grad.item(), want.grad.item(), rtol=2e-5)
    # first-order LPT: no second set of gradients
    d1, v1 = _Displace.apply(disp0, vel0, fac[0].detach(), fac[1].detach(), None, None, *[t.detach() for t in g[:3]])
    assert torch.equal(d1, torch.stack([disp0[:, i] + f32[0].detach() * g[i].detach() for i in range(3)], dim=-1))


@pytest.mark.parametrize('nranks, rank', [(2, 0), (2, 1), (4, 3), (8, 5)])
def test_slab_owner_matches_tensor_formula(nranks, rank):
    import pmwd_b200 as pm
    from pmwd_b200 import _lib, migrate
    conf = pm.Configuration(0.7, (32, 8, 8), mesh_shape=2)
    Mx = conf.mesh_shape[0]
    mx = Mx // nranks
    ptcl = pm.Particles.gen_grid(conf)
    disp = (ptcl.disp + 9.0 * conf.cell_size * _rand(ptcl.disp.shape, 7)).contiguous()
    n = disp.shape[0]
    owner = torch.empty(n, dtype=torch.uint8, device='cuda')
    need = torch.empty(1, dtype=torch.int32, device='cuda')
    _lib.check(_lib.lib().pmwd_slab_owner(_lib.stream_ptr(), n, _lib.ptr(ptcl.pmid), _lib.ptr(disp), float(conf.cell_size),
                                          Mx, nranks, rank * mx, mx, _lib.ptr(owner), _lib.ptr(need)), 'pmwd_slab_owner')
    want = migrate.owner_rank(ptcl.pmid[:, 0], disp[:, 0], conf, nranks)
    assert torch.equal(owner.to(torch.int64), want)
    cell32 = float(np.float32(conf.cell_size))
    plane = ptcl.pmid[:, 0].to(torch.int32) + torch.floor(disp[:, 0] / cell32).to(torch.int32)
    d = torch.remainder(plane - rank * mx, Mx)
    right = d + (2 - mx)
    ref = torch.where(d < mx, right.clamp(min=0), torch.minimum(right, Mx - d)).max().item()
    assert need.item() == ref


@pytest.mark.parametrize('ty, bw', [(0, 0), (8, 16)])
def test_two_source_sort_and_permute(ty, bw):
    """pmwd_cell_sort_perm2 + pmwd_permute_rows2 (the slab re-sort with Eulerian ownership: old rows minus the
    ones that left, plus the arrivals, read as one virtual array) == compacting by hand and sorting one array."""
    import pmwd_b200 as pm
    from pmwd_b200 import _lib
    from pmwd_b200.gravity import _force_desc
    lib, st = _lib.lib(), _lib.stream_ptr()
    conf = pm.Configuration(1., (16, 16, 16), mesh_shape=2)
    ptcl = pm.Particles.gen_grid(conf)
    disp = (ptcl.disp + 2.5 * conf.cell_size * _rand(ptcl.disp.shape, 11)).contiguous()
    vel = _rand(ptcl.disp.shape, 12)
    n_all = disp.shape[0]
    nA = 3000
    g = torch.Generator(device='cuda').manual_seed(13)
    owner = (torch.rand(nA, device='cuda', generator=g) < 0.2).to(torch.uint8)      # 1 = left for "rank 1"
    A = dict(pmid=ptcl.pmid[:nA].contiguous(), disp=disp[:nA].contiguous(), vel=vel[:nA].contiguous())
    B = dict(pmid=ptcl.pmid[nA:].contiguous(), disp=disp[nA:].contiguous(), vel=vel[nA:].contiguous())
    nB = n_all - nA
    keep = owner == 0
    n = int(keep.sum()) + nB

    def sort(desc, pmidA, dispA, na, own, pmidB, dispB):
        perm = torch.empty(desc.ptcl_num, dtype=torch.int32, device='cuda')
        scratch = torch.empty(lib.pmwd_cell_sort_scratch_bytes(C.byref(desc)), dtype=torch.uint8, device='cuda')
        _lib.check(lib.pmwd_cell_sort_perm2(st, C.byref(desc), _lib.ptr(pmidA), _lib.ptr(dispA), na, _lib.ptr(own), 0,
                                            _lib.ptr(pmidB), _lib.ptr(dispB), _lib.ptr(perm), _lib.ptr(scratch),
                                            scratch.numel(), ty, bw, None, None, 0.0), 'pmwd_cell_sort_perm2')
        return perm

    desc = _force_desc(ptcl.pmid, conf)
    desc.ptcl_num = n_all
    perm = sort(desc, A['pmid'], A['disp'], nA, owner, B['pmid'], B['disp'])
    names = ('pmid', 'disp', 'vel')
    got = {k: torch.empty((n,) + tuple(A[k].shape[1:]), dtype=A[k].dtype, device='cuda') for k in names}
    vp = C.c_void_p * 3
    rb = (C.c_int32 * 3)(6, 12, 12)
    _lib.check(lib.pmwd_permute_rows2(st, n, _lib.ptr(perm), 3, vp(*[A[k].data_ptr() for k in names]), nA,
                                      vp(*[B[k].data_ptr() for k in names]), vp(*[got[k].data_ptr() for k in names]), rb),
               'pmwd_permute_rows2')
    # by hand: compact, concatenate, single-source stable sort
    cat = {k: torch.cat([A[k][keep], B[k]]).contiguous() for k in names}
    desc1 = _force_desc(cat['pmid'], conf)
    perm1 = sort(desc1, cat['pmid'], cat['disp'], n, None, None, None)
    torch.cuda.synchronize()
    for k in names:
        assert torch.equal(got[k], cat[k][perm1.long()]), k
    # the rows that left are the tail of the permutation
    tail = perm[n:].long()
    assert tail.numel() == nA - int(keep.sum()) and bool((tail < nA).all()) and bool((owner[tail] == 1).all())


def test_reciprocal_cell_path_is_bitwise_the_division(monkeypatch):
    """Power-of-two cells: x * (1 / cell) replaces the IEEE division x / cell (pm_util.py:133, gather.py:108-110)
    in the per-particle CIC kernels; both are the correctly rounded value of the same real number, so forcing
    the division (PMWD_CIC_DIV=1) must not change a single bit of acc or of the displacement cotangent."""
    import pmwd_b200 as pm
    from pmwd_b200.gravity import force_adj_into
    n = 32
    conf = pm.Configuration(1., (n, n, n), mesh_shape=2, scatter_mode='deterministic')
    ptcl = pm.Particles.gen_grid(conf)
    disp = (ptcl.disp + 2.3 * _rand(ptcl.disp.shape, 21)).contiguous()
    pi = _rand(ptcl.disp.shape, 22)
    out = []
    for force_div in (False, True):
        if force_div:
            monkeypatch.setenv('PMWD_CIC_DIV', '1')
        acc, alpha = torch.empty_like(disp), torch.empty_like(disp)
        force_adj_into(ptcl.pmid, disp, 0.3, conf, pi, acc, alpha)
        torch.cuda.synchronize()
        out.append((acc, alpha))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert out[0][1].abs().max().item() > 0
