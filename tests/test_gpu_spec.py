"""GPU parity of the power-spectrum estimator (pmwd_b200.powspec, csrc/powspec.cu) against the
oracle (pmwd/spec_util.py:50-147).  Tolerances: mode counts exact, <k> 1e-12, P(k) 1e-5 relative
(float32 FFTs of two libraries; the north star asks for 0.1 %).

The per-mode arithmetic is also held to the oracle on the CPU (tests/test_powspec_host.py)."""

import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


def _check(got, want):
    k, P, N, b = (t.cpu().numpy() for t in got)
    np.testing.assert_array_equal(N, want[2])
    np.testing.assert_allclose(b, want[3], rtol=1e-14)
    ok = want[2] > 0
    np.testing.assert_allclose(k[ok], want[0][ok], rtol=1e-12)
    np.testing.assert_allclose(P[ok], want[1][ok], rtol=1e-5, atol=1e-6 * np.abs(want[1][ok]).max())


@pytest.mark.parametrize('shape', [(16, 16, 16), (12, 10, 9), (64, 64, 64), (40, 48, 33)])
@pytest.mark.parametrize('bins, deconv, cross', [(1j / 3, None, False), (1, 2, False), (1j / 2, None, True),
                                                 ((0, 0.1, 0.25, 0.4, 0.7), 1, False)])
def test_powspec_vs_oracle(shape, bins, deconv, cross):
    import pmwd_b200 as pm
    rng = np.random.default_rng(5)
    f = rng.standard_normal(shape).astype(np.float32)
    g = (0.5 * f + rng.standard_normal(shape)).astype(np.float32) if cross else None
    for cut_zero, cut_nyq in ((True, True), (False, False)):
        want = O.powspec(f, 0.7, bins=bins, g=g, deconv=deconv, cut_zero=cut_zero, cut_nyq=cut_nyq)
        got = pm.powspec(torch.from_numpy(f).cuda(), 0.7, bins=bins,
                         g=None if g is None else torch.from_numpy(g).cuda(), deconv=deconv,
                         cut_zero=cut_zero, cut_nyq=cut_nyq)
        _check(got, want)


def test_powspec_sums_leading_axes_and_rejects_bad_input():
    import pmwd_b200 as pm
    rng = np.random.default_rng(6)
    f = rng.standard_normal((3, 16, 16, 16)).astype(np.float32)
    _check(pm.powspec(torch.from_numpy(f).cuda(), 1.0), O.powspec(f, 1.0))
    with pytest.raises(pm._lib.PmwdError):
        pm.powspec(torch.from_numpy(f), 1.0)                     # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        pm.powspec(torch.from_numpy(f).cuda(), 1.0, g=torch.zeros(2, 2, 2, device='cuda'))
    with pytest.raises(ValueError):
        pm.powspec(torch.from_numpy(f).cuda(), 1.0, bins=(0.1, 0.2))


def test_powspec_of_evolved_density_256():
    """P(k) of a 256^3 density mesh (scatter of displaced particles): the parity metric at a size
    where the binning kernel runs many rows per warp."""
    import pmwd_b200 as pm
    n = 128
    conf = pm.Configuration(1., (n, n, n), mesh_shape=2)
    ptcl = pm.Particles.gen_grid(conf)
    g = torch.Generator(device='cuda').manual_seed(0)
    ptcl = ptcl.replace(disp=ptcl.disp + 2.0 * torch.randn(ptcl.disp.shape, device='cuda', generator=g))
    dens = pm.scatter(ptcl, conf)
    _check(pm.powspec(dens, conf.cell_size), O.powspec(dens.cpu().numpy(), conf.cell_size))


@pytest.mark.parametrize('shape, bins, deconv', [((16, 16, 16), 1j / 3, None), ((12, 10, 9), 1, 2), ((32, 24, 20), 1j / 2, 1)])
def test_powspec_gradient_vs_torch_float64(shape, bins, deconv):
    """d/df of sum_b c_b P_b through pmwd_b200.powspec's autograd against a plain float64 torch
    restatement of the estimator (rfftn + bucketize + index_add) differentiated by torch."""
    import math
    import pmwd_b200 as pm
    from pmwd_b200.spec_util import _getbins
    rng = np.random.default_rng(8)
    f = rng.standard_normal(shape).astype(np.float32)
    spacing = 0.7
    bnum, bcut, edges, right = _getbins(shape, bins, True)

    x = torch.from_numpy(f).cuda().requires_grad_(True)
    k, P, N, _ = pm.powspec(x, spacing, bins=bins, deconv=deconv)
    c = torch.from_numpy(rng.standard_normal(P.shape[0])).cuda()
    ok = N > 0
    (torch.where(ok, c * P, torch.zeros_like(P))).sum().backward()

    y = torch.from_numpy(f).double().cuda().requires_grad_(True)
    fk = torch.fft.rfftn(y)
    Pk = fk.real ** 2 + fk.imag ** 2
    ks = [torch.fft.fftfreq(n, dtype=torch.float64).float() for n in shape[:-1]] + \
         [torch.fft.rfftfreq(shape[-1], dtype=torch.float64).float()]
    ks = [kk.cuda().reshape([-1 if a == i else 1 for a in range(3)]) for i, kk in enumerate(ks)]
    kabs = torch.sqrt((ks[0] ** 2 + ks[1] ** 2) + ks[2] ** 2).expand(Pk.shape)
    if deconv is not None:
        for kk in ks:
            Pk = Pk * (torch.sinc(kk.double()) ** -deconv)
    mult = torch.full(Pk.shape, 2.0, dtype=torch.float64, device='cuda')
    mult[..., 0] = 1
    if shape[-1] % 2 == 0:
        mult[..., -1] = 1
    b = torch.bucketize(kabs.double().contiguous(), torch.tensor(edges, dtype=torch.float64, device='cuda'),
                        right=not right)
    sums = torch.zeros(bnum + 1, dtype=torch.float64, device='cuda').index_add(0, b.reshape(-1), (Pk * mult).reshape(-1))
    cnt = torch.zeros(bnum + 1, dtype=torch.float64, device='cuda').index_add(0, b.reshape(-1), mult.reshape(-1))
    Pref = sums[1:bcut] / cnt[1:bcut] * (spacing ** 3 / math.prod(shape))
    assert torch.equal(cnt[1:bcut], N)
    np.testing.assert_allclose(P.detach().cpu().numpy()[ok.cpu().numpy()], Pref.detach().cpu().numpy()[ok.cpu().numpy()],
                               rtol=1e-5)
    (torch.where(ok, c * Pref, torch.zeros_like(Pref))).sum().backward()
    g, gref = x.grad.double(), y.grad
    assert (g - gref).abs().max().item() <= 1e-4 * gref.abs().max().item()
