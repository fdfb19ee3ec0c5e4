#!/usr/bin/env python
"""bench.py -- PM steps/s and particle-updates/s, forward and forward+adjoint (driver contract).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 512]

A "step" is one kick-drift-force-kick leapfrog step (pmwd/nbody.py:121-140) over all
particles.  At N=1 the workload is BASELINE.json configs[2]'s geometry, 512^3 particles on a
1024^3 mesh (the size the metric is quoted on); the W+K steps are the first steps of the
default 63-step schedule a = 1/64 -> 1 starting from 2LPT initial conditions of a synthetic
Gaussian field (seed 0), so with the defaults (W=3, K=60) the timed region is the whole run.

`value`      whole-job forward particle-updates/s, state resident in HBM, CUDA events, max over ranks.
`fwd_adjoint`  the other half of BASELINE.json's metric: K forward steps + the reverse-time adjoint
             (nbody_adj: init + K adjoint steps) back over the same section of the schedule;
             `value` there = N_p * K / (t_forward + t_adjoint) = gradient particle-steps/s.
`e2e`        same metric through the public API (`nbody_step`) with HOST particle arrays:
             pinned-host -> device copies of the step's inputs and device -> host copies of its
             outputs are inside the timed region, every step.
`roofline`   the hand-written kernel FURTHEST below its roofline (lowest frac; deterministic choice):
             algorithmic bytes / CUDA-event time, vs the measured HBM copy bandwidth
             (MEASURED_PEAKS.json).  `kernels` lists every stage, forward and adjoint.
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
# NCCL_DEBUG stays as the caller set it: file descriptor 1 is pointed at stderr for the whole life of
# the GPU arm (NCCL prints its banner / INFO lines on stdout), the ONE JSON line goes to the real stdout.


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
def host_threads():
    """Cores this process may run on (torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm
    must not inherit that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port():
    """The CPU implementation that is timed: dict of step functions, description, threads used."""
    thr = host_threads()
    os.environ['OMP_NUM_THREADS'] = str(thr)         # read by libgomp when oracle/cpm.c's library loads
    import oracle as O
    try:
        from oracle import cpm
        cpm.lib()
        import functools
        fns = {k: functools.partial(getattr(cpm, k), threads=thr)
               for k in ('nbody_init', 'nbody_step', 'nbody_adj_init', 'nbody_adj_step')}
        return (fns, f'C + OpenMP port of pmwd (oracle/cpm.c: scatter with atomic adds, 3-mesh gather, k-space, '
                     f'kick/drift and their adjoints on {thr} threads; FFTs scipy pocketfft on all cores)', thr)
    except Exception as e:     # no gcc / libgomp on this box: fall back to the NumPy oracle, and say so
        print(f'[bench] oracle/cpm.c unavailable ({e}); timing the NumPy oracle', file=sys.stderr)
        import numpy as np

        def adj_init(a, ptcl, cot, cosmo, conf):
            ptcl, cot, ccf = O.nbody.force_adj(a, ptcl, cot, cosmo, conf)
            return ptcl, cot, {'Omega_m': np.float64(0), 'growth': np.zeros_like(cosmo.growth)}, ccf
        fns = {'nbody_init': O.nbody_init, 'nbody_step': O.nbody_step, 'nbody_adj_init': adj_init,
               'nbody_adj_step': lambda a0, a1, p, c, cosmo, cc, ccf, conf: O.integrate_adj(
                   a0, a1, p, c, cosmo, cc, ccf, conf)}
        return (fns, 'NumPy oracle port of pmwd (FFTs on all cores via scipy pocketfft, CIC loops '
                     'single-threaded NumPy)', 1)


def cpu_steps(n, steps, warmup, adjoint=True):
    """Time `steps` forward KDK steps of the CPU port at n^3 particles / (2n)^3 mesh, then (adjoint)
    the reverse-time adjoint back over the same steps (init + `steps` adjoint steps)."""
    fns, desc, thr = cpu_port()
    import numpy as np
    import oracle as O
    conf = O.Conf(1., (n, n, n), mesh_shape=2)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    modes = O.linear_modes(O.white_noise(0, conf), cosmo, conf)
    ptcl = O.lpt(modes, cosmo, conf)
    a = conf.a_nbody
    nsched = len(a) - 1
    ptcl = fns['nbody_init'](a[0], ptcl, cosmo, conf)
    i = 0
    for _ in range(warmup):
        ptcl = fns['nbody_step'](a[i % nsched], a[i % nsched + 1], ptcl, cosmo, conf); i += 1
    i0 = i
    t0 = time.perf_counter()
    for _ in range(steps):
        ptcl = fns['nbody_step'](a[i % nsched], a[i % nsched + 1], ptcl, cosmo, conf); i += 1
    dt = time.perf_counter() - t0
    assert np.isfinite(ptcl['disp']).all()
    out = {'value': conf.ptcl_num * steps / dt, 's_per_step': dt / steps, 'desc': desc, 'threads': thr}
    if adjoint:
        rng = np.random.default_rng(1)
        cot = {'disp': rng.standard_normal(ptcl['disp'].shape, dtype=np.float32),
               'vel': rng.standard_normal(ptcl['disp'].shape, dtype=np.float32),
               'acc': np.zeros_like(ptcl['disp'])}
        t0 = time.perf_counter()
        ptcl, cot, cc, ccf = fns['nbody_adj_init'](a[(i - 1) % nsched + 1], ptcl, cot, cosmo, conf)
        for j in range(i - 1, i0 - 1, -1):
            ptcl, cot, cc, ccf = fns['nbody_adj_step'](a[j % nsched + 1], a[j % nsched], ptcl, cot, cosmo, cc, ccf,
                                                       conf)
        dta = time.perf_counter() - t0
        assert np.isfinite(cot['disp']).all()
        out['adj_s_per_step'] = dta / steps
        out['fwd_adjoint_value'] = conf.ptcl_num * steps / (dt + dta)
    return out


def pick_cpu_sample(steps, warmup, budget_s=200.):
    # measured s per forward step of the C port on 8 cores; an adjoint step costs ~2.5x a forward one
    cost = {256: 3.6, 128: 0.9, 64: 0.2, 32: 0.05}
    for n in (256, 128, 64, 32):
        if (steps * 3.5 + warmup) * cost[n] <= budget_s:
            return n
    return 32


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.cpu_n if args.cpu_n_given else pick_cpu_sample(args.steps, args.warmup)
    r = cpu_steps(n, args.steps, args.warmup)
    sample = (f'{n}^3 particles / {2 * n}^3 mesh (bounded sample of the {args.n}^3/{2 * args.n}^3 workload; '
              f'per-particle rates), {args.steps} forward + {args.steps} adjoint steps; ' + r['desc'])
    cfg = workload_config(args, args.gpus)
    cfg['cpu_sample'] = {'ptcl_grid': [n] * 3, 'mesh': [2 * n] * 3, 'threads': r['threads']}
    line = {
        'impl': 'reference', 'metric': 'particle_updates_per_sec', 'value': r['value'],
        'unit': 'particle-updates/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': r['s_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': cfg,
        'cpu_baseline': {'value': r['value'], 'unit': 'particle-updates/s', 'cores': r['threads'], 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': r['value'], 'unit': 'particle-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'fwd_adjoint': {'value': r['fwd_adjoint_value'], 'unit': 'particle-steps/s (one forward + one adjoint step)',
                        'ms_per_step': (r['s_per_step'] + r['adj_s_per_step']) * 1e3,
                        'adjoint_ms_per_step': r['adj_s_per_step'] * 1e3},
        'steps_per_sec': 1.0 / r['s_per_step'],
    }
    print(json.dumps(line), flush=True)


def workload_config(args, ngpu):
    shape = rank_grid(args.n, ngpu)
    return {'workload': f'KDK PM steps (forward; forward + reverse-time adjoint in fwd_adjoint), '
                        f'{shape[0]}x{shape[1]}x{shape[2]} particles, '
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
_REAL_STDOUT = None


def redirect_stdout_to_stderr():
    """Point fd 1 at stderr for the rest of the process (NCCL with NCCL_DEBUG set, and other libraries,
    print on stdout at any time, also at teardown); keep the real stdout for the ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    """Print the JSON line on the real stdout."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


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
    # ---- reverse-time adjoint (the "fwd+adjoint" half of the metric): nbody_adj back over the section
    # of the schedule the forward leg just timed (init + K adjoint steps), CUDA events
    adjoint = None
    fwd_adjoint = None
    alg_adj = {'kick_drift_adj': 108 * Np, 'memset': 4 * Nm, 'scatter': 18 * Np + 4 * Nm,
               'scatter3': 30 * Np + 12 * Nm, 'fft_r2c': 8 * Nm, 'fft_c2r': 8 * Nm, 'kspace_force': 16 * Nm,
               'kspace_force_adj': 16 * Nm, 'gather3': 54 * Np + 12 * Nm,
               # weight-gradient gather + the acc gather of the same force meshes in one pass
               'force_adj_gather': 54 * Np + 16 * Nm}
    if not args.no_adjoint:
        i_end = stepper.i
        ka = max(1, min(K, i_end))
        section = a[i_end - ka:i_end + 1]
        with torch.no_grad():
            disp, vel = store.lagrangian('disp', 'vel')
            final = pm.Particles(conf, ic.pmid, disp, vel=vel)
            g = torch.Generator(device=dev).manual_seed(1)
            cot = pm.Particles(conf, ic.pmid, torch.randn(disp.shape, device=dev, generator=g),
                               vel=torch.randn(disp.shape, device=dev, generator=g))
            del store, stepper, disp, vel
            torch.cuda.empty_cache()
            # warm-up (untimed): the adjoint over the last 3 steps of the section -- first launches of the
            # adjoint kernels and, above all, the caching allocator's first device allocations of the ~25 GB of
            # particle / cotangent / workspace buffers (200 ms of cudaMalloc inside the timed region otherwise:
            # 114 vs 124 ms per step pair between two identical runs, gpurun r2x)
            pm.nbody_adj(final, cot, None, cosmo, conf, _a_nbody=section[-min(4, len(section)):])
            torch.cuda.synchronize()
            _lib.profile_enable(True); _lib.profile_read()
            la0 = _lib.launch_count()
            torch.cuda.synchronize()
            e0.record()
            _, pc, cc = pm.nbody_adj(final, cot, None, cosmo, conf, _a_nbody=section)
            e1.record()
            torch.cuda.synchronize()
            ams = e0.elapsed_time(e1)
            launches_adj = _lib.launch_count() - la0
            astages = _lib.profile_read()
            _lib.profile_enable(False)
        assert torch.isfinite(pc.disp).all() and torch.isfinite(pc.vel).all()
        aper = ams / ka                      # the init force_adj is charged to the steps (conservative)
        akern = {}
        for name, (tms, calls) in astages.items():
            if calls == 0:
                continue
            ent = {'ms_per_step': round(tms / ka, 3), 'launches': calls, 'share_of_step': round(tms / ams, 4)}
            if name in alg_adj:
                per = tms / calls
                ach = alg_adj[name] / per / 1e6
                ent.update(ms_per_launch=round(per, 4), alg_bytes=alg_adj[name], achieved_GBps=round(ach, 1),
                           frac=round(ach / peak, 4))
            akern[name] = ent
        adjoint = {'ms_per_step': aper, 'steps': ka, 'includes': 'nbody_adj init (one force_adj) + steps',
                   'particle_steps_per_sec': Np / (aper * 1e-3),
                   'adjoint_over_forward': aper / (ms / K),
                   'alg_bytes_per_step': 312 * Np + 156 * Nm,
                   'step_frac': round((312 * Np + 156 * Nm) / aper / 1e6 / peak, 4),
                   'gpu_launches': launches_adj, 'kernels': akern}
        fwd_adjoint = {'value': Np / ((aper + ms / K) * 1e-3),
                       'unit': 'particle-steps/s (one forward + one adjoint step)',
                       'ms_per_step': aper + ms / K, 'adjoint_ms_per_step': aper, 'forward_ms_per_step': ms / K,
                       'alg_bytes_per_step': 444 * Np + 224 * Nm,
                       'step_frac': round((444 * Np + 224 * Nm) / (aper + ms / K) / 1e6 / peak, 4)}
        del final, cot, pc
        torch.cuda.empty_cache()
    store = stepper = None

    # ---- roofline headline: the hand-written kernel with the LOWEST fraction of its roofline, forward
    # or adjoint (a deterministic choice: every stage is listed in `kernels` / `adjoint.kernels`)
    cand = [('forward', k, v) for k, v in kernels.items()
            if 'frac' in v and not k.startswith('fft_') and k != 'memset']
    if adjoint:
        cand += [('adjoint', k, v) for k, v in adjoint['kernels'].items()
                 if 'frac' in v and not k.startswith('fft_') and k != 'memset' and k not in kernels]
    leg, dom, dv = min(cand, key=lambda t: t[2]['frac'])
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    roofline = {'kernel': dom, 'leg': leg, 'bound': 'hbm', 'achieved': dv['achieved_GBps'], 'peak': peak,
                'unit': 'GB/s', 'frac': dv['frac'], 'traffic': traffic, 'peak_source': peak_src,
                'selection': 'lowest frac among the hand-written kernels (forward and adjoint)',
                'step_alg_bytes': 132 * Np + 68 * Nm,
                'step_frac': round((132 * Np + 68 * Nm) / (ms / K) / 1e6 / peak, 4)}

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
    # nbody_step_host keeps pmid (constant during a run) resident after its first upload, and does not
    # re-upload the accelerations it wrote itself on the previous step (the device keeps that array)
    h2d = sum(host[k].numel() * host[k].element_size() for k in ('disp', 'vel'))
    h2d_plain = sum(host[k].numel() * host[k].element_size() for k in ('pmid', 'disp', 'vel', 'acc'))
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
            api = ('pmwd_b200.nbody_step_host: pinned host arrays in and out every step (disp, vel up; disp, vel, acc '
                   'down); acc is not re-uploaded (device mirror of the array the previous step wrote); disp download '
                   'under the force, acc download under the next step\'s uploads (full-duplex link); checked against '
                   'nbody_step + explicit copies on the warm-up step')
        else:
            print('[bench] nbody_step_host differs from the plain route; timing the plain route', file=sys.stderr)
    except Exception as e:                         # noqa: BLE001 -- the plain route is always available
        print(f'[bench] nbody_step_host unavailable ({e!r}); timing the plain route', file=sys.stderr)
        torch.cuda.synchronize()
        e2e_plain(0, host, host)
    del chk
    e2e_step(1 % nsched, host, host)               # second warm-up: from here on acc is the array the API wrote
    torch.cuda.synchronize()
    e0.record()
    for j in range(2, 2 + ke):
        e2e_step(j % nsched, host, host)
    e1.record()
    torch.cuda.synchronize()
    ems = e0.elapsed_time(e1)
    e2e = {'value': Np * ke / (ems * 1e-3), 'unit': 'particle-updates/s', 'steps': ke,
           'ms_per_step': ems / ke, 'h2d_bytes_per_step': h2d if e2e_step is e2e_host else h2d_plain,
           'd2h_bytes_per_step': d2h, 'api': api}
    del host
    pm.nbody_host_release()
    torch.cuda.empty_cache()

    # ---- CPU baseline on a bounded sample (rank 0, N=1)
    cpu = None
    cfg = workload_config(args, 1)
    if not args.no_cpu_baseline:
        ncpu = args.cpu_n
        r = cpu_steps(ncpu, 3, 1, adjoint=not args.no_adjoint)
        cpu = {'value': r['value'], 'unit': 'particle-updates/s', 'cores': r['threads'], 'kind': 'port',
               'sample': f'{ncpu}^3 particles / {2 * ncpu}^3 mesh, 3 forward KDK steps after 1 warm-up '
                         f'({r["s_per_step"]:.2f} s/step)'
                         + (f' + the adjoint back over them ({r["adj_s_per_step"]:.2f} s/step)'
                            if 'adj_s_per_step' in r else '') + '; ' + r['desc']}
        if 'fwd_adjoint_value' in r:
            cpu['fwd_adjoint_value'] = r['fwd_adjoint_value']
        cfg['cpu_sample'] = {'ptcl_grid': [ncpu] * 3, 'mesh': [2 * ncpu] * 3, 'threads': r['threads']}

    # ---- context: a literal torch-CUDA transcription of the reference step (index_put_, advanced-index
    # gather, torch.fft) on THIS GPU, bounded size -- stands in for "the reference's JAX-on-GPU path"
    context = {'h100_pcie_jax_derived_updates_per_s': 6.5e8,
               'note': 'BASELINE.md derived figure for the same geometry on other hardware; not a published '
                       'number for this metric, hence vs_baseline = null'}
    if not args.no_context:
        try:
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            import torch_baseline
            torch.cuda.empty_cache()
            rate, tms = torch_baseline.time_steps(256, 3, dev)
            context['torch_cuda_transcription_updates_per_s'] = rate
            context['torch_cuda_transcription'] = (f'tools/torch_baseline.py, 256^3 particles / 512^3 mesh, forward '
                                                   f'KDK step, {tms:.1f} ms/step on this GPU')
            torch.cuda.empty_cache()
        except Exception as e:                     # noqa: BLE001 -- context only
            context['torch_cuda_transcription_error'] = repr(e)

    line = {
        'metric': 'particle_updates_per_sec', 'value': value, 'unit': 'particle-updates/s',
        'n_gpus': 1, 'steps': K, 'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': cfg,
        'steps_per_sec': K / (ms * 1e-3),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'storage_reorders': reorders,
        'roofline': roofline, 'kernels': kernels, 'fwd_adjoint': fwd_adjoint, 'adjoint': adjoint,
        'cpu_baseline': cpu, 'context': context,
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
    ap.add_argument('--cpu-n', type=int, default=None,
                    help='CPU sample size per side (default: 256 for the in-line cpu_baseline; the '
                         'reference arm picks the largest size that fits its time budget)')
    ap.add_argument('--no-context', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-adjoint', action='store_true')
    args = ap.parse_args()
    args.cpu_n_given = args.cpu_n is not None
    if args.cpu_n is None:
        args.cpu_n = 256
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        redirect_stdout_to_stderr()
        run_ours(args)


if __name__ == '__main__':
    main()
