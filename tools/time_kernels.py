"""Per-kernel CUDA-event timings of the hot path (development aid; bench.py is the contract).

usage: python tools/time_kernels.py [--n 256 512] [--sigma 0.2 2 10] [--out gpurun_out/time.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmwd_b200 as pm  # noqa: E402
from pmwd_b200 import _lib  # noqa: E402
from pmwd_b200.gravity import force_into, force_adj_into, _force_desc  # noqa: E402

HBM = 6552.0  # GB/s, MEASURED_PEAKS.json


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def run(n, sigma, results):
    conf = pm.Configuration(1., (n, n, n), mesh_shape=2)
    lib = _lib.lib()
    ptcl = pm.Particles.gen_grid(conf, vel=True, acc=True)
    g = torch.Generator(device='cuda').manual_seed(0)
    disp = ptcl.disp + sigma * conf.cell_size * torch.randn(ptcl.disp.shape, device='cuda', generator=g)
    vel = torch.randn(ptcl.disp.shape, device='cuda', generator=g)
    ptcl = ptcl.replace(disp=disp.contiguous(), vel=vel)
    Np, Nm = conf.ptcl_num, conf.mesh_size
    desc = _force_desc(ptcl.pmid, conf)
    st = _lib.stream_ptr()
    mesh = torch.zeros(conf.mesh_shape, device='cuda')
    F = [torch.randn(conf.mesh_shape, device='cuda', generator=g) for _ in range(3)]
    spec_shape = conf.mesh_shape[:-1] + (conf.mesh_shape[-1] // 2 + 1,)
    rho_k = torch.randn(spec_shape, dtype=torch.complex64, device='cuda')
    gk = [torch.empty_like(rho_k) for _ in range(3)]
    ctx = _lib.Context.get(torch.device('cuda', torch.cuda.current_device())).reserve(conf.mesh_shape)
    shp = _lib.shape_arr(conf.mesh_shape)
    row = {}

    def rec(name, ms, nbytes):
        row[name] = dict(ms=round(ms, 4), GBps=round(nbytes / ms / 1e6, 1), frac=round(nbytes / ms / 1e6 / HBM, 3))

    def scat(mode, scratch=None, sb=0):
        mesh.zero_()
        _lib.check(lib.pmwd_scatter(st, C.byref(desc), _lib.ptr(ptcl.pmid), _lib.ptr(ptcl.disp), None, 8.0,
                                    _lib.ptr(mesh), mode, _lib.ptr(scratch), sb), 'scatter')
    rec('memset+scatter_atomic', timeit(lambda: scat(0)), 18 * Np + 8 * Nm)
    sb = lib.pmwd_scatter_scratch_bytes(C.byref(desc), 1)
    scratch = torch.empty(sb, dtype=torch.uint8, device='cuda')
    rec('memset+scatter_det', timeit(lambda: scat(1, scratch, sb)), 18 * Np + 8 * Nm)
    del scratch
    rec('memset', timeit(lambda: mesh.zero_()), 4 * Nm)

    acc = ptcl.acc

    def gath3():          # pmwd_gather3: the three force meshes in one pass (gravity.py:61-70)
        _lib.check(lib.pmwd_gather3(st, C.byref(desc), _lib.ptr(ptcl.pmid), _lib.ptr(ptcl.disp), _lib.ptr(F[0]),
                                    _lib.ptr(F[1]), _lib.ptr(F[2]), _lib.ptr(acc), None, 0.0), 'gather3')
    rec('gather3', timeit(gath3), 30 * Np + 12 * Nm)

    def kforce():
        arr = (C.c_void_p * 3)(*[g_.data_ptr() for g_ in gk])
        _lib.check(lib.pmwd_kspace_force(st, 3, shp, conf.cell_size, 1.0, _lib.ptr(rho_k), arr), 'ks')
    rec('kspace_force', timeit(kforce), 16 * Nm)

    def r2c():
        _lib.check(lib.pmwd_fft_r2c(ctx.handle, st, 3, shp, _lib.ptr(F[0]), _lib.ptr(rho_k)), 'r2c')
    rec('cufft_r2c', timeit(r2c), 8 * Nm)

    def c2r():
        _lib.check(lib.pmwd_fft_c2r(ctx.handle, st, 3, shp, _lib.ptr(gk[0]), _lib.ptr(F[1]), 1.0), 'c2r')
    rec('cufft_c2r', timeit(c2r), 8 * Nm)

    def kd():
        _lib.check(lib.pmwd_kick_drift(st, 3 * Np, _lib.ptr(ptcl.disp), _lib.ptr(ptcl.vel), _lib.ptr(acc),
                                       1e-9, 1e-9, 1, 1), 'kd')
    rec('kick_drift', timeit(kd), 60 * Np)

    rec('force_fwd', timeit(lambda: force_into(ptcl.pmid, ptcl.disp, 0.3, conf, acc)), 72 * Np + 68 * Nm)
    rec('force_fwd+kick', timeit(lambda: force_into(ptcl.pmid, ptcl.disp, 0.3, conf, acc, ptcl.vel, 1e-9)),
        72 * Np + 68 * Nm)
    pi = torch.randn_like(ptcl.disp)
    alpha = torch.empty_like(pi)
    rec('force_adj', timeit(lambda: force_adj_into(ptcl.pmid, ptcl.disp, 0.3, conf, pi, acc, alpha), reps=3),
        (72 + 30 + 42) * Np + (68 + 24 + 24 + 16 + 8 + 16) * Nm)
    from pmwd_b200.nbody import _Store
    store = _Store(conf, dict(pmid=ptcl.pmid, disp=ptcl.disp, vel=ptcl.vel, acc=ptcl.acc))
    store.reorder()
    rec('reorder(sort+permute 4 arrays)', timeit(store.reorder, reps=3, warm=1), 2 * 46 * Np)
    sp = store.ptcl
    rec('force_fwd_after_sort', timeit(lambda: force_into(sp.pmid, sp.disp, 0.3, conf, sp.acc)), 72 * Np + 68 * Nm)
    step = row['kick_drift']['ms'] + row['force_fwd+kick']['ms']
    row['fwd_step'] = dict(ms=round(step, 3), updates_per_s=round(Np / step * 1e3, 0),
                           frac=round((132 * Np + 68 * Nm) / step / 1e6 / HBM, 3))
    results[f'n{n}_sigma{sigma}'] = row
    print(f'n{n}_sigma{sigma}', json.dumps(row), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, nargs='+', default=[256])
    ap.add_argument('--sigma', type=float, nargs='+', default=[0.2, 10.0])
    ap.add_argument('--out', default='gpurun_out/time.json')
    a = ap.parse_args()
    res = {}
    for n in a.n:
        for s in a.sigma:
            run(n, s, res)
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, 'w'), indent=1)
