#!/usr/bin/env python
"""bench.py -- PM steps/s and particle-updates/s of the forward KDK step (driver contract).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 512]

A "step" is one kick-drift-force-kick leapfrog step (pmwd/nbody.py:121-140) over all
particles.  At N=1 the workload is BASELINE.json configs[2]'s geometry, 512^3 particles on a
1024^3 mesh (the size the metric is quoted on); the W+K steps are the first steps of the
default 63-step schedule a = 1/64 -> 1 starting from 2LPT initial conditions of a synthetic
Gaussian field (seed 0), so with the defaults (W=3, K=60) the timed region is the whole run.

`value`      whole-job particle-updates/s, state resident in HBM, CUDA events, max over ranks.
`e2e`        same metric through the public API (`nbody_step`) with HOST particle arrays:
             pinned-host -> device copies of the step's inputs and device -> host copies of its
             outputs are inside the timed region, every step.
`roofline`   dominant hand-written kernel: algorithmic bytes / CUDA-event time, vs the measured
             HBM copy bandwidth (MEASURED_PEAKS.json).  `kernels` lists every stage.
`cpu_baseline`  the oracle's compiled port of the reference algorithm (oracle/cpm.c: C + OpenMP
             CIC / k-space / kick-drift loops, scipy pocketfft FFTs, all host cores) on a bounded
             sample of the same workload; the NumPy oracle if the C library cannot be built.
`--impl reference`  times that CPU implementation as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.pop('NCCL_DEBUG', None)       # NCCL_DEBUG >= VERSION prints a banner on stdout: ONE JSON line only


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                 '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------- CPU oracle arm
def cpu_port():
    """The CPU implementation that is timed: (step functions, description, threads used)."""
    import oracle as O
    try:
        from oracle import cpm
        cpm.lib()
        thr = cpm.max_threads()
        return (cpm.nbody_init, cpm.nbody_step,
                f'C + OpenMP port of pmwd (oracle/cpm.c: scatter with atomic adds, 3-mesh gather, k-space, '
                f'kick/drift on {thr} threads; FFTs scipy pocketfft on all cores)', thr)
    except Exception as e:     # no gcc / libgomp on this box: fall back to the NumPy oracle, and say so
        print(f'[bench] oracle/cpm.c unavailable ({e}); timing the NumPy oracle', file=sys.stderr)
        return (O.nbody_init, O.nbody_step,
                'NumPy oracle port of pmwd (FFTs on all cores via scipy pocketfft, CIC loops '
                'single-threaded NumPy)', 1)


def cpu_steps(n, steps, warmup):
    """Time `steps` forward KDK steps of the CPU port at n^3 particles / (2n)^3 mesh."""
    import numpy as np
    import oracle as O
    init, step, desc, thr = cpu_port()
    conf = O.Conf(1., (n, n, n), mesh_shape=2)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    modes = O.linear_modes(O.white_noise(0, conf), cosmo, conf)
    ptcl = O.lpt(modes, cosmo, conf)
    a = conf.a_nbody
    ptcl = init(a[0], ptcl, cosmo, conf)
    i = 0
    for _ in range(warmup):
        ptcl = step(a[i], a[i + 1], ptcl, cosmo, conf); i += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        ptcl = step(a[i % 63], a[i % 63 + 1], ptcl, cosmo, conf); i += 1
    dt = time.perf_counter() - t0
    assert np.isfinite(ptcl['disp']).all()
    return conf.ptcl_num * steps / dt, dt / steps, desc, thr


def pick_cpu_sample(steps, warmup, budget_s=150.):
    cost = {256: 3.6, 128: 0.9, 64: 0.2, 32: 0.05}   # measured s/step of the C port (8 cores)
    for n in (256, 128, 64, 32):
        if (steps + warmup) * cost[n] <= budget_s:
            return n
    return 32


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = pick_cpu_sample(args.steps, args.warmup)
    value, spstep, desc, cores = cpu_steps(n, args.steps, args.warmup)
    cores = max(cores, os.cpu_count() or 1)      # pocketfft runs on every core
    sample = (f'{n}^3 particles / {2 * n}^3 mesh (bounded sample of the {args.n}^3/{2 * args.n}^3 workload); '
              + desc)
    line = {
        'impl': 'reference', 'metric': 'particle_updates_per_sec', 'value': value,
        'unit': 'particle-updates/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': spstep * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'particle-updates/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'particle-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'steps_per_sec': 1.0 / spstep,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, ngpu):
    shape = rank_grid(args.n, ngpu)
    return {'workload': f'forward KDK PM steps, {shape[0]}x{shape[1]}x{shape[2]} particles, '
                        f'{2 * shape[0]}x{2 * shape[1]}x{2 * shape[2]} mesh (2x per side), 2LPT ICs at a=1/64, '
                        f'first W+K steps of the 63-step schedule to a=1',
            'ptcl_grid': list(shape), 'mesh': [2 * s for s in shape], 'ptcl_spacing_mpc_h': 1.0,
            'cosmology': 'SimpleLCDM', 'seed': 0, 'scatter_mode': args.scatter_mode,
            'reorder_every': args.reorder_every,
            'cache': 'inputs (>= 5 GB particle state, 4 GB meshes) far exceed the 126 MB L2; no flush needed',
            'parallelism': f'slab{ngpu}' if ngpu > 1 else 'single'}


def rank_grid(n, ngpu):
    """Weak scaling: per-GPU work fixed; the global box doubles along z, then y, then x (so
    that 8 GPUs run the 1024^3-particle / 2048^3-mesh configuration of BASELINE.json)."""
    shape = [n, n, n]
    g, ax = ngpu, 2
    while g > 1:
        shape[ax % 3] *= 2
        g //= 2
        ax -= 1
    return tuple(shape)


# ----------------------------------------------------------------------------- GPU arm
class _StdoutToStderr:
    """Route fd 1 to stderr while libraries initialise (NCCL prints a version banner on
    stdout), so that the only thing on stdout is the ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None

    def __exit__(self, *exc):
        self.restore()


def emit(line):
    """Print the JSON line on the real stdout."""
    guard = globals().get('_GUARD')
    if guard is not None:
        guard.restore()
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import pmwd_b200 as pm
    from pmwd_b200 import _lib
    from pmwd_b200.nbody import _Stepper, _store_from

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run for --gpus > 1')
    if world > 1:
        from pmwd_b200 import dist as pdist
        args._emit = emit
        return pdist.run_bench(args)

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    n = args.n
    conf = pm.Configuration(1., (n, n, n), mesh_shape=2, scatter_mode=args.scatter_mode, device=dev,
                            reorder_every=args.reorder_every, reorder_min_disp=args.reorder_min_disp)
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    with torch.no_grad():
        modes = pm.linear_modes(pm.white_noise(0, conf), cosmo, conf)
        ic, _ = pm.lpt(modes, cosmo, conf)
        del modes
    torch.cuda.empty_cache()
    a = conf.a_nbody.tolist()
    nsched = len(a) - 1
    Np, Nm = conf.ptcl_num, conf.mesh_size

    def fresh():
        sp = _Stepper(_store_from(ic, conf), a, cosmo, conf)
        sp.init()
        return sp

    W, K = args.warmup, args.steps
    with torch.no_grad():
        stepper = fresh()
        for _ in range(W):
            if stepper.i == nsched:
                stepper = fresh()
            stepper.step()
        torch.cuda.synchronize()
        sampler = ClockSampler(local); sampler.start()
        _lib.profile_enable(True); _lib.profile_read()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K):
            if stepper.i == nsched:
                stepper = fresh()
            stepper.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - l0
        stages = _lib.profile_read()
        _lib.profile_enable(False)
        clocks = sampler.stop()
    store = stepper.store
    assert torch.isfinite(store.arrays['disp']).all()
    reorders = store.reorders
    value = Np * K / (ms * 1e-3)

    # ---- per-stage roofline: algorithmic bytes per launch (SURVEY.md 8d / DESIGN.md)
    peak, peak_src = _peaks()
    alg = {'kick_drift': 60 * Np, 'memset': 4 * Nm, 'scatter': 18 * Np + 4 * Nm, 'fft_r2c': 8 * Nm,
           'kspace_force': 16 * Nm, 'fft_c2r': 8 * Nm,
           # gather + trailing half-kick + next step's half-kick and drift (pmwd_force_kdk)
           'gather3': 66 * Np + 12 * Nm}
    kernels = {}
    for name, (tms, calls) in stages.items():
        if calls == 0:
            continue
        if name not in alg:      # e.g. 'other' = cell sort + permutation of the particle storage
            kernels[name] = {'ms_total': round(tms, 3), 'launches': calls, 'share_of_step': round(tms / ms, 4)}
            continue
        per = tms / calls
        ach = alg[name] / per / 1e6
        kernels[name] = {'ms_per_launch': round(per, 4), 'launches': calls, 'share_of_step': round(tms / ms, 4),
                         'alg_bytes': alg[name], 'achieved_GBps': round(ach, 1), 'frac': round(ach / peak, 4)}
    ours = [k for k in kernels if k in alg and not k.startswith('fft_') and k != 'memset']
    dom = max(ours, key=lambda k: kernels[k]['share_of_step'])
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': kernels[dom]['achieved_GBps'], 'peak': peak,
                'unit': 'GB/s', 'frac': kernels[dom]['frac'], 'traffic': traffic, 'peak_source': peak_src,
                'step_alg_bytes': 132 * Np + 68 * Nm,
                'step_frac': round((132 * Np + 68 * Nm) / (ms / K) / 1e6 / peak, 4)}

    # ---- reverse-time adjoint (the "fwd+adjoint" half of the metric): the whole nbody_adj
    # loop from the evolved state (63 adjoint steps + init), CUDA events
    adjoint = None
    if not args.no_adjoint:
        with torch.no_grad():
            disp, vel = store.lagrangian('disp', 'vel')
            final = pm.Particles(conf, ic.pmid, disp, vel=vel)
            g = torch.Generator(device=dev).manual_seed(1)
            cot = pm.Particles(conf, ic.pmid, torch.randn(disp.shape, device=dev, generator=g),
                               vel=torch.randn(disp.shape, device=dev, generator=g))
            del store, stepper
            torch.cuda.empty_cache()
            _lib.profile_enable(True); _lib.profile_read()
            torch.cuda.synchronize()
            e0.record()
            _, pc, cc = pm.nbody_adj(final, cot, None, cosmo, conf)
            e1.record()
            torch.cuda.synchronize()
            ams = e0.elapsed_time(e1)
            astages = _lib.profile_read()
            _lib.profile_enable(False)
        assert torch.isfinite(pc.disp).all()
        nadj = nsched
        adjoint = {'ms_per_step': ams / nadj, 'steps': nadj,
                   'particle_steps_per_sec': Np * nadj / (ams * 1e-3),
                   'adjoint_over_forward': (ams / nadj) / (ms / K),
                   'gradient_particle_steps_per_sec': Np / ((ams / nadj + ms / K) * 1e-3),
                   'alg_bytes_per_step': 312 * Np + 156 * Nm,
                   'step_frac': round((312 * Np + 156 * Nm) / (ams / nadj) / 1e6 / peak, 4),
                   'stage_ms_per_step': {k: round(v[0] / nadj, 3) for k, v in astages.items() if v[1]}}
        del final, cot, pc
        store = None
        torch.cuda.empty_cache()

    # ---- e2e: public API with host buffers, copies inside the timed region
    ke = max(1, min(K, args.e2e_steps))
    host = {k: torch.empty(getattr(ic, k).shape, dtype=getattr(ic, k).dtype).pin_memory()
            for k in ('pmid', 'disp', 'vel')}
    host['acc'] = torch.empty(ic.disp.shape, dtype=ic.disp.dtype).pin_memory()
    p0, _ = pm.nbody_init(a[0], ic, None, cosmo, conf)
    for k in host:
        host[k].copy_(getattr(p0, k))
    del p0
    torch.cuda.empty_cache()
    h2d = sum(t.numel() * t.element_size() for t in host.values())
    d2h = sum(host[k].numel() * host[k].element_size() for k in ('disp', 'vel', 'acc'))

    def e2e_plain(j, src, dst):
        d = {k: src[k].to(dev, non_blocking=True) for k in src}
        p = pm.Particles(conf, d['pmid'], d['disp'], vel=d['vel'], acc=d['acc'])
        p, _ = pm.nbody_step(a[j], a[j + 1], p, None, cosmo, conf)
        for k in ('disp', 'vel', 'acc'):
            dst[k].copy_(getattr(p, k), non_blocking=True)

    def e2e_host(j, src, dst):
        pm.nbody_step_host(a[j], a[j + 1], src, cosmo, conf, out=dst)

    # The host-array entry point overlaps the displacement download with the force; it is used if,
    # on this box, it reproduces the plain route (copy up, nbody_step, copy down) on the warm-up step.
    api = 'pmwd_b200.nbody_step on pinned host arrays (explicit copies around it)'
    e2e_step = e2e_plain
    chk = None
    try:
        chk = {k: torch.empty_like(host[k]).pin_memory() for k in ('disp', 'vel', 'acc')}
        e2e_host(0, host, chk)
        torch.cuda.synchronize()
        e2e_plain(0, host, host)                   # warm-up (state rolls forward on the host)
        torch.cuda.synchronize()
        same = torch.equal(chk['disp'], host['disp'])
        for k in ('vel', 'acc'):
            scale = host[k].abs().max().item()
            same = same and (chk[k] - host[k]).abs().max().item() <= 1e-4 * scale
        if same:
            e2e_step = e2e_host
            api = ('pmwd_b200.nbody_step_host: pinned host arrays in and out, displacement download '
                   'overlapped with the force (checked against nbody_step + explicit copies on the warm-up step)')
        else:
            print('[bench] nbody_step_host differs from the plain route; timing the plain route', file=sys.stderr)
    except Exception as e:                         # noqa: BLE001 -- the plain route is always available
        print(f'[bench] nbody_step_host unavailable ({e!r}); timing the plain route', file=sys.stderr)
        torch.cuda.synchronize()
        e2e_plain(0, host, host)
    del chk
    torch.cuda.synchronize()
    e0.record()
    for j in range(1, 1 + ke):
        e2e_step(j % nsched, host, host)
    e1.record()
    torch.cuda.synchronize()
    ems = e0.elapsed_time(e1)
    e2e = {'value': Np * ke / (ems * 1e-3), 'unit': 'particle-updates/s', 'steps': ke,
           'ms_per_step': ems / ke, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'api': api}
    del host
    torch.cuda.empty_cache()

    # ---- CPU baseline on a bounded sample (rank 0, N=1)
    cpu = None
    if not args.no_cpu_baseline:
        ncpu = args.cpu_n
        v, sp, desc, thr = cpu_steps(ncpu, 3, 1)
        cpu = {'value': v, 'unit': 'particle-updates/s', 'cores': max(thr, os.cpu_count() or 1), 'kind': 'port',
               'sample': f'{ncpu}^3 particles / {2 * ncpu}^3 mesh, 3 forward KDK steps after 1 warm-up, '
                         f'{sp:.2f} s/step; ' + desc}

    line = {
        'metric': 'particle_updates_per_sec', 'value': value, 'unit': 'particle-updates/s',
        'n_gpus': 1, 'steps': K, 'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'steps_per_sec': K / (ms * 1e-3),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'storage_reorders': reorders,
        'roofline': roofline, 'kernels': kernels, 'adjoint': adjoint, 'cpu_baseline': cpu,
        'context': {'h100_pcie_jax_derived_updates_per_s': 6.5e8,
                    'note': 'BASELINE.md derived figure for the same geometry on other hardware; not a published '
                            'number for this metric, hence vs_baseline = null'},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n', type=int, default=512, help='particles per side per GPU (mesh is 2x)')
    ap.add_argument('--scatter-mode', default='atomic', choices=['atomic', 'deterministic'])
    ap.add_argument('--e2e-steps', type=int, default=4)
    ap.add_argument('--reorder-every', type=int, default=3)
    ap.add_argument('--reorder-min-disp', type=float, default=1.0)
    ap.add_argument('--cpu-n', type=int, default=256)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-adjoint', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        global _GUARD
        with _StdoutToStderr() as _GUARD:
            run_ours(args)


if __name__ == '__main__':
    main()
