"""A literal torch transcription of the reference's PM step, as the "framework baseline on the same GPU"
(SURVEY.md 8d): scatter = index_put_(accumulate=True) on materialised (N, 8) indices and weights,
gather = advanced indexing, FFTs = torch.fft, everything else elementwise -- i.e. what pmwd's JAX code
lowers to, without any of this repository's kernels.  NOT part of the product or of bench.py's arms;
it answers "how much of the speed-up is the hardware and how much the kernels".

usage: python tools/torch_baseline.py [n] [steps]     (n^3 particles, (2n)^3 mesh; CUDA if available)
With --check it runs 16^3 on the CPU against the oracle instead of timing.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def enmesh(pmid, disp, cell, shape):
    """pmwd/pm_util.py:119-138,154 (fast branch): indices (N, 8) linear int64, weights (N, 8)."""
    t = disp / cell
    i0 = torch.floor(t)
    nb = torch.arange(8, device=disp.device)
    bits = torch.stack([(nb >> a) & 1 for a in range(3)], dim=-1)                 # (8, 3)
    i = i0[:, None, :] + bits[None].to(disp.dtype)                                 # (N, 8, 3)
    w = (1 - (t[:, None, :] - i).abs())
    w = (w[..., 0] * w[..., 1]) * w[..., 2]
    idx = (pmid[:, None, :].long() + i.long()) % torch.tensor(shape, device=disp.device)
    lin = (idx[..., 0] * shape[1] + idx[..., 1]) * shape[2] + idx[..., 2]
    return lin, w


def scatter(pmid, disp, cell, shape, val):
    lin, w = enmesh(pmid, disp, cell, shape)
    mesh = torch.zeros(shape[0] * shape[1] * shape[2], dtype=disp.dtype, device=disp.device)
    mesh.index_put_((lin.reshape(-1),), (val * w).reshape(-1), accumulate=True)   # pmwd/scatter.py:80
    return mesh.reshape(shape)


def gather(pmid, disp, cell, mesh):
    lin, w = enmesh(pmid, disp, cell, mesh.shape)
    return (mesh.reshape(-1)[lin] * w).sum(dim=1)                                  # pmwd/gather.py:75


def gravity(pmid, disp, cell, shape, Omega_m, val):
    """pmwd/gravity.py:47-72."""
    dens = scatter(pmid, disp, cell, shape, val)
    dens = (dens - 1) * (1.5 * Omega_m)
    spec = torch.fft.rfftn(dens)
    ks = [torch.fft.fftfreq(n, device=disp.device, dtype=torch.float64) for n in shape[:-1]]
    ks.append(torch.fft.rfftfreq(shape[-1], device=disp.device, dtype=torch.float64))
    ks = [(k * (2 * np.pi / cell)).to(disp.dtype).reshape([-1 if a == i else 1 for a in range(3)])
          for i, k in enumerate(ks)]
    k2 = (ks[0] ** 2 + ks[1] ** 2) + ks[2] ** 2
    pot = torch.where(k2 != 0, -spec / torch.where(k2 != 0, k2, torch.ones_like(k2)), torch.zeros_like(spec))
    nyq = np.pi / cell
    eps = nyq * torch.finfo(disp.dtype).eps
    acc = []
    for k in ks:
        kk = torch.where((k.abs() - nyq).abs() <= eps, torch.zeros_like(k), k)
        F = torch.fft.irfftn(-1j * kk * pot, s=shape)
        acc.append(gather(pmid, disp, cell, F))
    return torch.stack(acc, dim=-1)


def step(pmid, disp, vel, acc, cell, shape, Omega_m, val, K1, D, K2):
    """KDK, pmwd/nbody.py:121-140 with the default splitting."""
    vel = vel + acc * K1
    disp = disp + vel * D
    acc = gravity(pmid, disp, cell, shape, Omega_m, val)
    vel = vel + acc * K2
    return disp, vel, acc


def check():
    import oracle as O
    n = 16
    conf = O.Conf(1., (n, n, n), mesh_shape=2)
    pmid, disp, _, _ = O.gen_grid(conf)
    disp = (disp + 1.5 * np.random.default_rng(0).standard_normal(disp.shape)).astype(np.float32)
    ref = O.gravity(pmid, disp, 0.3, conf)
    got = gravity(torch.from_numpy(pmid), torch.from_numpy(disp), conf.cell_size, conf.mesh_shape, 0.3,
                  conf.mesh_size / conf.ptcl_num).numpy()
    err = np.sqrt(np.mean((got - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
    print(f'torch transcription vs oracle at {n}^3: acc rel-RMS difference {err:.2e}')
    assert err < 1e-5
    return err


def time_steps(n=256, steps=5, dev=None):
    """(particle-updates/s, ms/step) of the transcription at n^3 particles / (2n)^3 mesh."""
    dev = dev or torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    shape = (2 * n,) * 3
    g = torch.Generator(device=dev).manual_seed(0)
    ax = torch.arange(n, device=dev, dtype=torch.int16) * 2
    pmid = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), dim=-1).reshape(-1, 3)
    disp = 2.0 * torch.randn((n ** 3, 3), device=dev, generator=g)
    vel = torch.zeros_like(disp)
    acc = gravity(pmid, disp, 0.5, shape, 0.3, 8.0)
    sync = torch.cuda.synchronize if dev.type == 'cuda' else (lambda: None)
    disp, vel, acc = step(pmid, disp, vel, acc, 0.5, shape, 0.3, 8.0, 1e-3, 1e-3, 1e-3)       # warm-up
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        disp, vel, acc = step(pmid, disp, vel, acc, 0.5, shape, 0.3, 8.0, 1e-3, 1e-3, 1e-3)
    sync()
    dt = (time.perf_counter() - t0) / steps
    return n ** 3 / dt, dt * 1e3


def main():
    if '--check' in sys.argv:
        check()
        return
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    n = int(args[0]) if args else 256
    steps = int(args[1]) if len(args) > 1 else 5
    dev = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    rate, ms = time_steps(n, steps, dev)
    print(f'torch transcription on {dev}: {n}^3 particles / {2 * n}^3 mesh, {ms:.1f} ms/step = '
          f'{rate:.3g} particle-updates/s')


if __name__ == '__main__':
    main()
