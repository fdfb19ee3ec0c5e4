"""N-body time integration and its reverse-time adjoint with the reference's API
(``pmwd/nbody.py:12-276``).

Per KDK step the GPU runs: one fused kick+drift pass, one fused force pipeline whose final
gather also applies the trailing half-kick (``pmwd_force``); the adjoint step runs one fused
kick_adj+drift_adj pass (state + cotangents + float64 dot products), ``pmwd_force_adj``
and one kick_adj pass.  Step factors are float64 host scalars (``nbody.py:12-36``) cast
to float32 before touching particles.  ``jax.custom_vjp`` becomes a
``torch.autograd.Function`` whose residual is only the final state (``nbody.py:263-265``).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .boltzmann import growth
from .cosmology import E2, H_deriv
from .gravity import gravity, force_into, force_adj_into, force_kdk_into, _force_desc
from .particles import Particles
from . import sweep as _sweep


def _G_D(a, cosmo, conf):
    """Growth factor of ZA canonical velocity in [H_0] (``nbody.py:12-14``)."""
    return a ** 2 * torch.sqrt(E2(a, cosmo)) * growth(a, cosmo, conf, deriv=1)


def _G_K(a, cosmo, conf):
    """Growth factor of ZA accelerations in [H_0^2] (``nbody.py:17-22``)."""
    return a ** 3 * E2(a, cosmo) * (
        growth(a, cosmo, conf, deriv=2)
        + (2 + H_deriv(a, cosmo)) * growth(a, cosmo, conf, deriv=1))


def drift_factor(a_vel, a_prev, a_next, cosmo, conf):
    """Drift time step factor in [1/H_0] (``nbody.py:25-29``)."""
    factor = growth(a_next, cosmo, conf) - growth(a_prev, cosmo, conf)
    return factor / _G_D(a_vel, cosmo, conf)


def kick_factor(a_acc, a_prev, a_next, cosmo, conf):
    """Kick time step factor in [1/H_0] (``nbody.py:32-36``)."""
    factor = _G_D(a_next, cosmo, conf) - _G_D(a_prev, cosmo, conf)
    return factor / _G_K(a_acc, cosmo, conf)


def _f(a):
    return float(a)


def _f32(x):
    """``factor.astype(conf.float_dtype)`` (``nbody.py:42,73``)."""
    return float(np.float32(float(x)))


# ------------------------------------------------------------------------------- kernels
def _kick_drift(ptcl, K, D, do_kick, do_drift):
    dev = ptcl.disp.device
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pmwd_kick_drift(
            _lib.stream_ptr(dev), ptcl.disp.numel(), _lib.ptr(ptcl.disp), _lib.ptr(ptcl.vel),
            _lib.ptr(ptcl.acc), K, D, int(do_kick), int(do_drift)), 'pmwd_kick_drift')


def _owned(ptcl, conf, need_acc=True):
    """Private, contiguous, writable copies of the dynamic arrays (inputs are immutable)."""
    _lib.require_cuda(ptcl.pmid, ptcl.disp)
    if ptcl.vel is None:
        raise ValueError('nbody needs particle velocities')
    kw = dict(pmid=ptcl.pmid.contiguous(),
              disp=ptcl.disp.detach().clone(memory_format=torch.contiguous_format),
              vel=ptcl.vel.detach().clone(memory_format=torch.contiguous_format))
    if need_acc:
        kw['acc'] = (ptcl.acc.detach().clone(memory_format=torch.contiguous_format)
                     if ptcl.acc is not None else torch.zeros_like(kw['disp']))
    return Particles(conf, attr=ptcl.attr, **kw)


def _fast_ok(ptcl, conf):
    return conf.dim == 3 and ptcl.pmid.dtype == torch.int16


class _Store:
    """The integrator's private particle arrays, optionally kept in mesh-cell order.

    ``arrays`` maps name -> (N, 3) tensor (``pmid`` int16, the rest float32).  ``reorder()``
    re-sorts all of them by the particles' current mesh cell (``pmwd_cell_sort_perm`` +
    ``pmwd_permute_rows``); ``lagrangian(name)`` returns an array restored to the reference's
    Lagrangian order.  Per-particle arithmetic does not depend on the storage order."""

    def __init__(self, conf, arrays):
        self.conf = conf
        self.arrays = dict(arrays)
        self.lag = None            # uint32 view as int32 tensor: Lagrangian index of each slot
        self._alt = None           # spare buffer set of the re-sort (name -> tensor with >= N rows)
        self._back = None          # the full-capacity tensors behind `arrays` / `lag` when they are views
        self._perm = None
        self._scratch = None
        self.steps_since = 0
        self.active = False
        self.reorders = 0
        self.desc_fn = None        # pmid -> pmwd_cic_desc used for the sort keys (slab runs)
        self.sweep = None          # sweep.SweepState: table + scratch of the tiled deposit
        self.migrator = None       # dist.SlabComm: particles move to the rank that owns their current x plane
        self.migrations = 0
        self.slab_sweep = False    # slab runs: keep the storage in the tiled deposit's order as well
        self.det_sweep = False     # deterministic mode through the reproducible tiled deposit (sort before every force)
        self.timers = None         # dist.TIMERS in slab runs: 'migrate' / 'resort' phases of the bench line

    def _timed(self, name):
        import contextlib
        return self.timers(name) if self.timers is not None else contextlib.nullcontext()

    def setup_sweep(self):
        """Tiled deposit (csrc/scatter_sweep.cu) for the single-device fast path: the storage is sorted
        into the tile order right away (one sort per run) and the table rebuilt from every re-sort."""
        conf, a = self.conf, self.arrays
        if self.desc_fn is not None or not _sweep.enabled(conf) or conf.dim != 3 \
                or a['pmid'].dtype != torch.int16 or conf.reorder_every <= 0:
            return
        desc = _force_desc(a['pmid'], conf)
        det = conf.scatter_mode == 'deterministic'
        st = _sweep.SweepState(desc, a['disp'].device, det=det)
        if st.ty > 0 and (st.det or not det):
            self.sweep = st
            self.det_sweep = det
            self.reorder()

    def sweep_arg(self):
        return self.sweep.arg() if self.sweep is not None else None

    def sort_for_force(self):
        """Deterministic mode: the reproducible tiled deposit needs storage sorted from the very positions it
        deposits, so the re-sort happens right before every force evaluation (and never in maybe_reorder)."""
        if self.det_sweep:
            self.reorder()

    def check_det(self):
        """End of a run in deterministic mode (one synchronisation): no deposit may have met a particle outside
        its tile's window -- it would have gone through order-dependent REDs."""
        if self.det_sweep and self.sweep is not None:
            bad = self.sweep.det_violations()
            if bad:
                raise _lib.PmwdError(f'deterministic deposit: {bad} particle deposits went through the atomic '
                                     'straggler path (storage not re-sorted before a force evaluation)')

    @property
    def ptcl(self):
        a = self.arrays
        return Particles(self.conf, a['pmid'], a['disp'], vel=a['vel'], acc=a['acc'])

    def maybe_reorder(self, sync_max=None, predict=0.0):
        """Re-sort every ``conf.reorder_every`` steps.  ``predict``: sort by the PREDICTED positions
        ``disp + vel * predict`` (a drift factor), so that the order is centred on the force evaluations until
        the next re-sort instead of being exact for the first and two drifts stale for the last
        (``PMWD_SORT_PREDICT=0`` turns it off)."""
        conf = self.conf
        if conf.reorder_every <= 0 or conf.dim != 3 or self.arrays['pmid'].dtype != torch.int16 or self.det_sweep:
            return False
        self.steps_since += 1
        if self.steps_since < conf.reorder_every:
            return False
        self.steps_since = 0
        if not self.active:
            # one cheap reduction + sync every `reorder_every` steps until structure has formed
            m = self.arrays['disp'].abs().max()
            m = sync_max(m) if sync_max is not None else float(m)
            if m < conf.reorder_min_disp * conf.cell_size:
                return False
            self.active = True
        if os.environ.get('PMWD_SORT_PREDICT', '1') == '0':
            predict = 0.0
        self.reorder(predict)
        return True

    def reorder(self, predict=0.0):
        moved = None
        if self.migrator is not None:
            with self._timed('migrate'):
                moved = self._migrate()
        with self._timed('resort'):
            self._resort(moved, predict)

    def _migrate(self):
        """Eulerian ownership (slab runs): every particle moves to the rank that owns its current base plane.
        Only the movers travel (pmwd_b200/migrate.py); the particles that stay are NOT compacted here -- the
        re-sort that follows reads the old arrays and the arrivals as one virtual array and drops the rows
        that left (``pmwd_cell_sort_perm2`` / ``pmwd_permute_rows2``).  `lag` carries the GLOBAL Lagrangian
        index so that lagrangian() can send everything home again."""
        from . import migrate
        a, comm = self.arrays, self.migrator
        dev = a['disp'].device
        n = a['disp'].shape[0]
        if self.lag is None:
            self.lag = comm.rank * n + torch.arange(n, dtype=torch.int32, device=dev)
        owner, arrivals, nmove = migrate.exchange_movers(dict(a, lag=self.lag), self.conf, comm.group)
        self.migrations += 1
        return owner, arrivals, nmove

    def _resort(self, moved=None, predict=0.0):
        conf, a = self.conf, self.arrays
        dev = a['disp'].device
        nA = a['disp'].shape[0]
        lib = _lib.lib()
        if self.lag is None:
            self.lag = torch.arange(nA, dtype=torch.int32, device=dev)   # bit pattern of uint32
        owner, arrivals, nmove = moved if moved is not None else (None, None, 0)
        nB = arrivals['disp'].shape[0] if arrivals is not None else 0
        ntot, n = nA + nB, nA + nB - nmove          # rows to sort, rows of the new storage
        desc = self.desc_fn(a['pmid']) if self.desc_fn is not None else _force_desc(a['pmid'], conf)
        desc.ptcl_num = n
        if self.slab_sweep and self.desc_fn is not None and _sweep.enabled(conf):
            # slab + halo planes: the table follows the descriptor of THIS re-sort (a later change of the
            # halo width makes it unusable until the next one; the RED kernel takes over meanwhile)
            if self.sweep is None:
                st = _sweep.SweepState(desc, dev)
                self.sweep = st if st.ty > 0 else None
                self.slab_sweep = self.sweep is not None
            else:
                self.sweep.ok = False
                self.sweep.fit(desc)
        cur = dict(a, lag=self.lag)
        names = list(cur)
        # the spare buffer set: capacity with some room when the row count changes with every migration
        # (exact-size buffers would mean fresh device allocations at every re-sort)
        alt = self._alt
        if alt is None or set(alt) != set(names) or any(alt[k].shape[0] < n for k in names):
            cap = n if self.migrator is None else n + n // 16
            alt = {k: torch.empty((cap,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev) for k, v in cur.items()}
        if self._perm is None or self._perm.numel() < ntot:
            self._perm = torch.empty(ntot if self.migrator is None else ntot + ntot // 16, dtype=torch.int32, device=dev)
        # the sort's scratch follows the row count and the mesh the descriptor spans: re-query every time
        sdesc = self.desc_fn(a['pmid']) if self.desc_fn is not None else _force_desc(a['pmid'], conf)
        sdesc.ptcl_num = ntot
        nbytes = lib.pmwd_cell_sort_scratch_bytes(C.byref(sdesc))
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty(nbytes + (nbytes // 16 if self.migrator is not None else 0), dtype=torch.uint8,
                                        device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            ty, bw = (self.sweep.ty, self.sweep.bw) if self.sweep is not None else (0, 0)
            rank = self.migrator.rank if self.migrator is not None else 0
            _lib.check(lib.pmwd_cell_sort_perm2(
                st, C.byref(sdesc), _lib.ptr(a['pmid']), _lib.ptr(a['disp']), nA, _lib.ptr(owner), rank,
                _lib.ptr(arrivals['pmid']) if nB else None, _lib.ptr(arrivals['disp']) if nB else None,
                _lib.ptr(self._perm), _lib.ptr(self._scratch), self._scratch.numel(), ty, bw,
                _lib.ptr(a['vel']) if predict else None, _lib.ptr(arrivals['vel']) if (predict and nB) else None,
                float(predict)), 'pmwd_cell_sort_perm2')
            vp = C.c_void_p * len(names)
            src = vp(*[cur[k].data_ptr() for k in names])
            srcB = vp(*[arrivals[k].data_ptr() for k in names]) if nB else None
            dst = vp(*[alt[k].data_ptr() for k in names])
            rb = (C.c_int32 * len(names))(*[cur[k].element_size() * (cur[k].shape[1] if cur[k].ndim > 1 else 1)
                                            for k in names])
            _lib.check(lib.pmwd_permute_rows2(st, n, _lib.ptr(self._perm), len(names), src, nA, srcB, dst, rb),
                       'pmwd_permute_rows2')
        # ping-pong: the buffers behind the old arrays become the spare set (checked for capacity next time)
        self._alt = self._back if self._back is not None else cur
        self._back = alt
        new = {k: alt[k][:n] for k in names}
        self.lag = new.pop('lag')
        self.arrays = new
        self.reorders += 1
        if self.sweep is not None:
            # table of the new order from the sort's own keys (valid by construction: no read-back)
            keys = lib.pmwd_cell_sort_sorted_keys(C.byref(sdesc), _lib.ptr(self._scratch))
            self.sweep.build(desc, C.c_void_p(keys))

    def lagrangian(self, *names):
        """Arrays restored to Lagrangian order (new tensors; storage untouched)."""
        a = self.arrays
        if self.migrator is not None and self.migrations > 0:
            # collective: every rank sends its particles back to their Lagrangian owner (sorted by `lag`)
            from . import migrate
            comm = self.migrator
            home = migrate.to_lagrangian(dict({k: a[k] for k in names}, lag=self.lag.to(torch.int64)),
                                         self.conf.ptcl_num, comm.group)
            outs = [home[k].contiguous() for k in names]
            return outs if len(outs) > 1 else outs[0]
        if self.lag is None:
            out = [a[k].clone() for k in names]
            return out if len(out) > 1 else out[0]
        dev = a['disp'].device
        n = a['disp'].shape[0]
        outs = [torch.empty_like(a[k]) for k in names]
        with torch.cuda.device(dev):
            src = (C.c_void_p * len(names))(*[a[k].data_ptr() for k in names])
            dst = (C.c_void_p * len(names))(*[o.data_ptr() for o in outs])
            rb = (C.c_int32 * len(names))(*[a[k].element_size() * a[k].shape[1] for k in names])
            _lib.check(_lib.lib().pmwd_permute_rows(_lib.stream_ptr(dev), n, _lib.ptr(self.lag), len(names),
                                                    src, dst, rb, 1), 'pmwd_permute_rows')
        return outs if len(outs) > 1 else outs[0]


def _store_from(ptcl, conf, _tiled=True, **extra):
    p = _owned(ptcl, conf)
    # pmid is cloned too: the store ping-pongs between two buffer sets when it re-sorts and
    # must never write into the caller's tensors
    pmid = p.pmid.clone() if conf.reorder_every > 0 else p.pmid
    arrays = dict(pmid=pmid, disp=p.disp, vel=p.vel, acc=p.acc)
    arrays.update(extra)
    store = _Store(conf, arrays)
    if _tiled:
        store.setup_sweep()
    return store


# --------------------------------------------------------------- reference-shaped pieces
def drift(a_vel, a_prev, a_next, ptcl, cosmo, conf):
    """``nbody.py:39-46``."""
    factor = _f32(drift_factor(a_vel, a_prev, a_next, cosmo, conf))
    return ptcl.replace(disp=ptcl.disp + ptcl.vel * factor)


def kick(a_acc, a_prev, a_next, ptcl, cosmo, conf):
    """``nbody.py:70-77``."""
    factor = _f32(kick_factor(a_acc, a_prev, a_next, cosmo, conf))
    return ptcl.replace(vel=ptcl.vel + ptcl.acc * factor)


def force(a, ptcl, cosmo, conf):
    """``nbody.py:102-105``."""
    with torch.no_grad():
        acc = gravity(a, ptcl, cosmo, conf)
    return ptcl.replace(acc=acc)


def _integrate_inplace(a_prev, a_next, ptcl, cosmo, conf):
    """``integrate`` (``nbody.py:121-140``) on owned arrays, with kernel fusion: a kick that
    is followed by a drift shares its pass; a kick that follows a force rides on the gather."""
    Om = float(cosmo.Omega_m)
    fast = _fast_ok(ptcl, conf)
    D = K = 0
    a_disp = a_vel = a_acc = a_prev
    pending_kick = None           # factor of a kick not yet applied (fused with next drift)
    for d, k in conf.symp_splits:
        if d != 0:
            D += d
            a_disp_next = a_prev * (1 - D) + a_next * D
            fd = _f32(drift_factor(a_vel, a_disp, a_disp_next, cosmo, conf))
            if pending_kick is not None:
                _kick_drift(ptcl, pending_kick, fd, True, True)
                pending_kick = None
            else:
                _kick_drift(ptcl, 0.0, fd, False, True)
            a_disp = a_disp_next
            a_acc = a_disp
            # force, fused with the following kick when there is one in this split
            if k != 0:
                K += k
                a_vel_next = a_prev * (1 - K) + a_next * K
                fk = _f32(kick_factor(a_acc, a_vel, a_vel_next, cosmo, conf))
                if fast:
                    force_into(ptcl.pmid, ptcl.disp, Om, conf, ptcl.acc, ptcl.vel, fk)
                else:
                    ptcl.acc.copy_(gravity(a_disp, ptcl, cosmo, conf))
                    _kick_drift(ptcl, fk, 0.0, True, False)
                a_vel = a_vel_next
            else:
                if fast:
                    force_into(ptcl.pmid, ptcl.disp, Om, conf, ptcl.acc)
                else:
                    ptcl.acc.copy_(gravity(a_disp, ptcl, cosmo, conf))
            continue
        if k != 0:
            K += k
            a_vel_next = a_prev * (1 - K) + a_next * K
            fk = _f32(kick_factor(a_acc, a_vel, a_vel_next, cosmo, conf))
            if pending_kick is not None:
                _kick_drift(ptcl, pending_kick, 0.0, True, False)
            pending_kick = fk
            a_vel = a_vel_next
    if pending_kick is not None:
        _kick_drift(ptcl, pending_kick, 0.0, True, False)
    return ptcl


def integrate(a_prev, a_next, ptcl, cosmo, conf):
    """Symplectic integration for one step (``nbody.py:121-140``); returns new Particles."""
    with torch.no_grad():
        return _integrate_inplace(_f(a_prev), _f(a_next), _owned(ptcl, conf), cosmo, conf)


def form(a_prev, a_next, ptcl, cosmo, conf):
    pass


def coevolve(a_prev, a_next, ptcl, cosmo, conf):
    return ptcl


def observe(a_prev, a_next, ptcl, obsvbl, cosmo, conf):
    return obsvbl


def nbody_init(a, ptcl, obsvbl, cosmo, conf):
    """``nbody.py:193-201``."""
    with torch.no_grad():
        ptcl = _owned(ptcl, conf)
        _force_inplace(ptcl, cosmo, conf)
    return ptcl, obsvbl


def _force_inplace(ptcl, cosmo, conf):
    if _fast_ok(ptcl, conf):
        force_into(ptcl.pmid, ptcl.disp, float(cosmo.Omega_m), conf, ptcl.acc)
    else:
        ptcl.acc.copy_(gravity(0., ptcl, cosmo, conf))


def nbody_step(a_prev, a_next, ptcl, obsvbl, cosmo, conf):
    """``nbody.py:204-212``."""
    ptcl = integrate(a_prev, a_next, ptcl, cosmo, conf)
    return ptcl, obsvbl


_pmid_resident = {}        # (host data_ptr, shape, version, device) -> device copy of a host pmid array


def _resident_pmid(host_pmid, dev):
    """pmid never changes during a run (``pmwd/particles.py:52-59``): a host pmid array that has been
    uploaded before (same storage, same version counter) is not uploaded again."""
    key = (host_pmid.data_ptr(), tuple(host_pmid.shape), host_pmid._version, str(dev))
    t = _pmid_resident.get(key)
    if t is None:
        if len(_pmid_resident) >= 4:
            _pmid_resident.clear()
        t = _pmid_resident[key] = host_pmid.to(dev, non_blocking=True)
    return t


class _HostMirror:
    """Device-side working set of ``nbody_step_host`` on one device: the three dynamic arrays, two
    copy streams and what is known about the host array that mirrors the device accelerations."""

    def __init__(self, dev, shape):
        self.shape = tuple(shape)
        self.disp = torch.empty(shape, dtype=torch.float32, device=dev)
        self.vel = torch.empty(shape, dtype=torch.float32, device=dev)
        self.acc = torch.empty(shape, dtype=torch.float32, device=dev)
        self.side = torch.cuda.Stream(dev)       # displacement download, under the force
        self.side2 = torch.cuda.Stream(dev)      # acceleration download, under the next call's uploads
        self.acc_tag = None                      # (data_ptr, shape, version) of the host array written from self.acc
        self.acc_event = None                    # end of that download
        self.trace = [] if os.environ.get('PMWD_HOST_TRACE') else None    # (label, event) per stage, tools/time_host_step.py

    def mark(self, label, stream):
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self.trace.append((label, ev))


_host_mirrors = {}


def nbody_host_release(dev=None):
    """Drop the device buffers ``nbody_step_host`` keeps between calls (all devices by default)."""
    for k in list(_host_mirrors):
        if dev is None or k == str(torch.device(dev)):
            del _host_mirrors[k]
    _pmid_resident.clear()


def nbody_step_host(a_prev, a_next, host, cosmo, conf, out=None, acc_resident=True, _slab=None):
    """``nbody_step`` for a state that lives in (pinned) HOST memory: ``host`` and the returned
    dict map ``'pmid'`` (int16), ``'disp'``, ``'vel'``, ``'acc'`` (float32) to CPU tensors of
    shape ``(N, 3)``; ``out`` may supply the output buffers (``pmid`` is passed through).

    Same arithmetic as ``nbody_step`` (``nbody.py:204-212``).  The copies are part of the call:

    * ``disp`` and ``vel`` go up every call; ``pmid`` never changes (``pmwd/particles.py:52-59``) and is
      uploaded once per host array;
    * ``acc`` is the force at the displacements the previous call produced: the device keeps the array
      it computed, and a host ``acc`` that is the very array the previous call wrote (same storage,
      unchanged torch version counter) is not sent back up.  Any other ``acc`` array is uploaded;
      ``acc_resident=False`` always uploads (needed if the host array was modified behind torch's back,
      e.g. through a NumPy view);
    * the new displacements -- final after the drift -- come back on a second stream WHILE the force is
      computed, the velocities after the fused force + trailing half-kick, and the accelerations after
      those on a third stream, so that their download shares the (full-duplex) link with the NEXT call's
      uploads instead of delaying them.

    Everything is enqueued asynchronously: ``torch.cuda.synchronize(device)`` before reading the outputs
    (synchronising the current stream covers ``disp`` and ``vel`` only).  Default KDK splitting on the 3-D
    fast path; other configurations take the plain copy -> ``integrate`` -> copy route.

    ``_slab`` (internal, ``dist.nbody_step_slab_host``): a ``dist.SlabForce``; the arrays are then one rank's
    Lagrangian x-slab of a slab-decomposed run and the force is the collective one.
    """
    dev = torch.device(conf.device if getattr(conf, 'device', None) is not None else 'cuda')
    if dev.type != 'cuda' or not torch.cuda.is_available():
        raise _lib.PmwdError('pmwd_b200 kernels need a CUDA device (there is no CPU fallback)')
    for k in ('pmid', 'disp', 'vel', 'acc'):
        if host[k].is_cuda:
            raise ValueError(f"host['{k}'] is a CUDA tensor; use nbody_step for device-resident state")
    if out is None:
        out = {k: torch.empty_like(host[k]).pin_memory() for k in ('disp', 'vel', 'acc')}
    out['pmid'] = host['pmid']
    a_prev, a_next = _f(a_prev), _f(a_next)
    with torch.no_grad(), torch.cuda.device(dev):
        dev = torch.device('cuda', torch.cuda.current_device())
        cur = torch.cuda.current_stream(dev)
        m = _host_mirrors.get(str(dev))
        if m is None or m.shape != tuple(host['disp'].shape):
            _host_mirrors.pop(str(dev), None)
            m = _host_mirrors[str(dev)] = _HostMirror(dev, host['disp'].shape)
        m.mark('start', cur)
        tag = (host['acc'].data_ptr(), tuple(host['acc'].shape), host['acc']._version)
        if not (acc_resident and m.acc_tag == tag):
            if m.acc_event is not None:
                cur.wait_event(m.acc_event)      # the previous call's acc download still reads m.acc
                m.acc_event = None
            m.acc.copy_(host['acc'], non_blocking=True)
        m.acc_tag = None
        m.vel.copy_(host['vel'], non_blocking=True)
        m.disp.copy_(host['disp'], non_blocking=True)
        m.mark('uploaded', cur)
        pmid = _resident_pmid(host['pmid'], dev)
        p = Particles(conf, pmid, m.disp, vel=m.vel, acc=m.acc)
        default = tuple(tuple(x) for x in conf.symp_splits) == ((0, 0.5), (1, 0.5))
        if _slab is not None and not (default and _fast_ok(p, conf)):
            raise NotImplementedError('the slab integrator implements the default KDK splitting on the 3-D fast path')
        if not (default and _fast_ok(p, conf)):
            if m.acc_event is not None:
                cur.wait_event(m.acc_event)
                m.acc_event = None
            q = _integrate_inplace(a_prev, a_next, p, cosmo, conf)
            for k in ('disp', 'vel', 'acc'):
                out[k].copy_(getattr(q, k), non_blocking=True)
            if q.acc is not m.acc:
                m.acc.copy_(q.acc)
            m.acc_tag = (out['acc'].data_ptr(), tuple(out['acc'].shape), out['acc']._version)
            return out
        am = a_prev * 0.5 + a_next * 0.5
        K1 = _f32(kick_factor(a_prev, a_prev, am, cosmo, conf))
        D = _f32(drift_factor(am, a_prev, a_next, cosmo, conf))
        K2 = _f32(kick_factor(a_next, am, a_next, cosmo, conf))
        _kick_drift(p, K1, D, True, True)
        m.side.wait_stream(cur)
        with torch.cuda.stream(m.side):
            out['disp'].copy_(m.disp, non_blocking=True)
            m.mark('disp_down', m.side)
        m.mark('kick_drift', cur)
        if m.acc_event is not None:
            cur.wait_event(m.acc_event)          # the force overwrites m.acc: the previous call's download of it
            m.acc_event = None                   # (running under this call's uploads) must have finished
        if _slab is not None:
            _slab.force(p.pmid, p.disp, float(cosmo.Omega_m), p.acc, p.vel, K2)
        else:
            force_into(p.pmid, p.disp, float(cosmo.Omega_m), conf, p.acc, p.vel, K2)
        m.mark('force', cur)
        out['vel'].copy_(m.vel, non_blocking=True)
        m.mark('vel_down', cur)
        m.side2.wait_stream(cur)                 # after the velocities: same direction, same link
        with torch.cuda.stream(m.side2):
            out['acc'].copy_(m.acc, non_blocking=True)
            m.acc_event = m.side2.record_event()
            m.mark('acc_down', m.side2)
        m.acc_tag = (out['acc'].data_ptr(), tuple(out['acc'].shape), out['acc']._version)
        cur.wait_stream(m.side)                  # the next call may overwrite m.disp / read out['disp']
    return out


class _Stepper:
    """Pipelined KDK stepping over a schedule of scale factors on a ``_Store``: with the default
    splitting on the 3-D fast path every step is ONE ``pmwd_force_kdk`` call (force + trailing
    half-kick + the next step's leading half-kick and drift in the gather pass); the very first
    step's leading kick+drift is a separate ``pmwd_kick_drift`` pass.  Arithmetic per particle
    is the same float32 sequence as ``integrate`` (``nbody.py:121-140``).  Other configurations
    fall back to ``_integrate_inplace`` step by step."""

    def __init__(self, store, a_list, cosmo, conf):
        self.store, self.a, self.cosmo, self.conf = store, list(a_list), cosmo, conf
        self.i = 0
        self.fused = (tuple(tuple(x) for x in conf.symp_splits) == ((0, 0.5), (1, 0.5))
                      and _fast_ok(store.ptcl, conf))
        self._factors = {}
        self.pre = False          # the current step's leading kick+drift has already been applied

    def factors(self, i):
        f = self._factors.get(i)
        if f is None:
            a0, a1 = self.a[i], self.a[i + 1]
            am = a0 * 0.5 + a1 * 0.5
            f = (_f32(kick_factor(a0, a0, am, self.cosmo, self.conf)),
                 _f32(drift_factor(am, a0, a1, self.cosmo, self.conf)),
                 _f32(kick_factor(a1, am, a1, self.cosmo, self.conf)))
            self._factors[i] = f
        return f

    @property
    def nsteps(self):
        return len(self.a) - 1

    def init(self):
        self.store.sort_for_force()
        p = self.store.ptcl
        if _fast_ok(p, self.conf):
            force_into(p.pmid, p.disp, float(self.cosmo.Omega_m), self.conf, p.acc, sweep=self.store.sweep_arg())
        else:
            _force_inplace(p, self.cosmo, self.conf)

    def step(self):
        i = self.i
        st = self.store
        if not self.fused:
            _integrate_inplace(self.a[i], self.a[i + 1], st.ptcl, self.cosmo, self.conf)
        else:
            k1, d, k2 = self.factors(i)
            if not self.pre:
                _kick_drift(st.ptcl, k1, d, True, True)
            st.sort_for_force()
            a = st.arrays
            Om = float(self.cosmo.Omega_m)
            if i + 1 < self.nsteps:
                k1n, dn, _ = self.factors(i + 1)
                force_kdk_into(a['pmid'], a['disp'], Om, self.conf, a['acc'], a['vel'], k2, k1n, dn,
                               sweep=st.sweep_arg())
                self.pre = True
            else:
                force_into(a['pmid'], a['disp'], Om, self.conf, a['acc'], a['vel'], k2, sweep=st.sweep_arg())
                self.pre = False
        self.i += 1
        st.maybe_reorder(predict=self.predict())

    def predict(self):
        """Drift to add to the positions the re-sort keys are computed from.  The pipelined step has already
        applied the next step's drift, so the forces until the next re-sort see the positions now, one drift and
        two drifts later (``reorder_every`` = 3): centre the order on them."""
        every = self.conf.reorder_every
        if not self.pre or every < 2 or self.i + 1 >= self.nsteps:
            return 0.0
        return 0.5 * (every - 1) * self.factors(self.i + 1)[1]


def _nbody_forward(ptcl, cosmo, conf, reverse):
    a_nbody = conf.a_nbody.tolist()
    if reverse:
        a_nbody = a_nbody[::-1]
    with torch.no_grad():
        store = _store_from(ptcl, conf)
        stepper = _Stepper(store, a_nbody, cosmo, conf)
        stepper.init()
        for _ in range(stepper.nsteps):
            stepper.step()
        disp, vel, acc = store.lagrangian('disp', 'vel', 'acc')
        store.check_det()
    return Particles(conf, ptcl.pmid, disp, vel=vel, acc=acc, attr=ptcl.attr)


# ------------------------------------------------------------------------------- adjoint
_COSMO_LEAVES = ('Omega_m', 'Omega_k_', 'w_0_', 'w_a_', 'growth')


def _factor_valgrad(fun, a0, a1, a2, cosmo, conf):
    """``value_and_grad(fun, argnums=3)`` (``nbody.py:51-52,82-83``) w.r.t. the cosmology
    leaves a step factor can depend on."""
    with torch.enable_grad():
        leaves = {n: getattr(cosmo, n).detach().clone().requires_grad_(True)
                  for n in _COSMO_LEAVES if getattr(cosmo, n) is not None}
        c = cosmo.replace(**leaves)
        val = fun(a0, a1, a2, c, conf)
        grads = torch.autograd.grad(val, list(leaves.values()), allow_unused=True)
    grads = {n: (g if g is not None else torch.zeros_like(leaves[n]))
             for n, g in zip(leaves, grads)}
    return float(val.detach()), grads


def nbody_adj(ptcl, ptcl_cot, obsvbl_cot, cosmo, conf, reverse=False, _slab=None, _a_nbody=None):
    """N-body time integration with adjoint equation (``nbody.py:226-260``).

    ``_slab`` (internal): a ``dist.SlabForce`` when the particles are one rank's slab of a
    multi-GPU run; forces then go through the slab pipeline and the float64 dot products are
    all-reduced once at the end.  ``_a_nbody`` (internal, bench): a section of the schedule
    (increasing scale factors) to integrate back over instead of ``conf.a_nbody``.

    Returns ``(ptcl, ptcl_cot, cosmo_cot)``; ``cosmo_cot`` is a dict leaf-name -> float64
    tensor over the leaves ``nbody`` can touch (all other leaves of the reference's
    cotangent pytree are zero).
    """
    if not _fast_ok(ptcl, conf):
        raise NotImplementedError('the reverse-time adjoint runs on the 3-D int16 fast path')
    a_nbody = conf.a_nbody.tolist() if _a_nbody is None else [float(x) for x in _a_nbody]
    if reverse:
        a_nbody = a_nbody[::-1]
    dev = ptcl.disp.device
    Om = float(cosmo.Omega_m)
    lib = _lib.lib()
    pmid_in = ptcl.pmid
    with torch.no_grad():
        xi = ptcl_cot.disp.detach().to(conf.float_dtype).clone(memory_format=torch.contiguous_format)
        store = _store_from(
            ptcl, conf, _tiled=_slab is None, xi=xi,
            pi=ptcl_cot.vel.detach().to(conf.float_dtype).clone(memory_format=torch.contiguous_format),
            alpha=torch.empty_like(xi))
        if _slab is not None:
            store.desc_fn = lambda pmid: _slab._desc(pmid, max(_slab.h_alloc, 1))
            import os
            if os.environ.get('PMWD_MIGRATE', '1') != '0' and _slab.comm.size > 1:
                store.migrator = _slab.comm
                store.slab_sweep = os.environ.get('PMWD_SLAB_SWEEP', '1') != '0'
                from .dist import TIMERS as _T
                store.timers = _T
        sync_max = _slab.comm.allreduce_max if _slab is not None else None
        m0 = store.arrays['disp'].abs().max()
        m0 = sync_max(m0) if sync_max is not None else float(m0)     # one decision for all ranks
        if conf.reorder_every > 0 and _fast_ok(ptcl, conf) and m0 >= conf.reorder_min_disp * conf.cell_size:
            store.active = True
            if store.reorders == 0:   # (the tiled deposit has already sorted the storage)
                store.reorder()       # the adjoint starts from the evolved (clustered) state
        pairs = list(zip(a_nbody[:0:-1], a_nbody[-2::-1]))
        # per kick / drift: float64 dot products kept on the device until the end
        nsplit = len(conf.symp_splits)
        sums = torch.zeros((len(pairs) * nsplit + 1, 2), dtype=torch.float64, device=dev)
        records = []   # (kind, slot, column, factor, grads)

        pending = []     # a trailing kick-only update waiting to share its pass with the next kick+drift
        last_drift = 0.0

        def _kd_call(K, D, do_kick, do_drift, slot):
            a = store.arrays
            with torch.cuda.device(dev):
                _lib.check(lib.pmwd_kick_drift_adj(
                    _lib.stream_ptr(dev), a['xi'].numel(), _lib.ptr(a['disp']), _lib.ptr(a['vel']),
                    _lib.ptr(a['acc']), _lib.ptr(a['xi']), _lib.ptr(a['pi']), _lib.ptr(a['alpha']), K, D,
                    int(do_kick), int(do_drift), C.c_void_p(sums[slot].data_ptr())),
                    'pmwd_kick_drift_adj')

        def flush_pending():
            if pending:
                K0, slot0 = pending.pop()
                _kd_call(K0, 0.0, True, False, slot0)

        def kd_adj(K, D, do_kick, do_drift, slot):
            if do_kick and not do_drift:
                flush_pending()
                pending.append((K, slot))      # acc / alpha stay as they are until the next force_adj
                return
            if pending and do_kick and do_drift:
                K0, slot0 = pending.pop()
                a = store.arrays
                with torch.cuda.device(dev):
                    _lib.check(lib.pmwd_kick_kick_drift_adj(
                        _lib.stream_ptr(dev), a['xi'].numel(), _lib.ptr(a['disp']), _lib.ptr(a['vel']), _lib.ptr(a['acc']),
                        _lib.ptr(a['xi']), _lib.ptr(a['pi']), _lib.ptr(a['alpha']), K0, K, D,
                        C.c_void_p(sums[slot0].data_ptr()), C.c_void_p(sums[slot].data_ptr())),
                        'pmwd_kick_kick_drift_adj')
                return
            flush_pending()
            _kd_call(K, D, do_kick, do_drift, slot)

        def f_adj():
            flush_pending()
            store.sort_for_force()
            a = store.arrays
            if _slab is not None:
                _slab.force_adj(a['pmid'], a['disp'], Om, a['pi'], a['acc'], a['alpha'], sweep=store.sweep)
            else:
                force_adj_into(a['pmid'], a['disp'], Om, conf, a['pi'], a['acc'], a['alpha'],
                               sweep=store.sweep_arg())

        # nbody_adj_init (nbody.py:226-236)
        f_adj()

        slot = 0
        for a_prev, a_next in pairs:
            # integrate_adj (nbody.py:143-162)
            K = D = 0
            a_disp = a_vel = a_acc = a_prev
            for d, k in reversed(conf.symp_splits):
                fk = fd = 0.0
                if k != 0:
                    K += k
                    a_vel_next = a_prev * (1 - K) + a_next * K
                    val, grads = _factor_valgrad(kick_factor, a_acc, a_vel, a_vel_next, cosmo, conf)
                    fk = _f32(val)
                    records.append(('kick', slot, 0, fk, grads))
                    a_vel = a_vel_next
                if d != 0:
                    D += d
                    a_disp_next = a_prev * (1 - D) + a_next * D
                    val, grads = _factor_valgrad(drift_factor, a_vel, a_disp, a_disp_next, cosmo, conf)
                    fd = _f32(val)
                    records.append(('drift', slot, 1, fd, grads))
                    a_disp = a_disp_next
                kd_adj(fk, fd, k != 0, d != 0, slot)
                slot += 1
                if d != 0:
                    f_adj()
                    a_acc = a_disp
                    last_drift = fd
            # the forces until the next re-sort come one, two and three drifts from here (reorder_every = 3)
            store.maybe_reorder(sync_max=sync_max, predict=0.5 * (conf.reorder_every + 1) * last_drift)

        flush_pending()
        if _slab is not None:
            _slab.comm.allreduce_sum_(sums)
        sums_h = sums.cpu()
        cosmo_cot = {nme: torch.zeros_like(getattr(cosmo, nme)) for nme in _COSMO_LEAVES
                     if getattr(cosmo, nme) is not None}
        for kind, s, col, factor, grads in records:
            S = sums_h[s, col]
            for nme in cosmo_cot:                        # nbody.py:64-65, 95-97
                cosmo_cot[nme] = cosmo_cot[nme] - grads[nme] * S
            if kind == 'kick':
                # cosmo_cot_force * factor: gravity's cosmology VJP has the single leaf
                # Omega_m, equal to sum(pi . acc) / Omega_m (gravity.py:54)
                cosmo_cot['Omega_m'] = cosmo_cot['Omega_m'] - (S / Om) * factor

        disp, vel, acc, xi, pi, alpha = store.lagrangian('disp', 'vel', 'acc', 'xi', 'pi', 'alpha')
        store.check_det()
    ptcl = Particles(conf, pmid_in, disp, vel=vel, acc=acc)
    ptcl_cot = Particles(conf, pmid_in, xi, vel=pi, acc=alpha)
    return ptcl, ptcl_cot, cosmo_cot


class _Nbody(torch.autograd.Function):
    """``nbody`` as a ``custom_vjp`` (``nbody.py:215-223,263-276``)."""

    @staticmethod
    def forward(ctx, disp, vel, conf, cosmo, pmid, reverse, *leaves):
        ptcl = Particles(conf, pmid, disp, vel=vel)
        out = _nbody_forward(ptcl, cosmo, conf, reverse)
        ctx.save_for_backward(out.disp, out.vel)      # residual = final state only
        ctx.meta = (conf, cosmo, out.pmid, reverse)
        ctx.mark_non_differentiable(out.acc)
        return out.disp, out.vel, out.acc

    @staticmethod
    def backward(ctx, disp_cot, vel_cot, _acc_cot):
        conf, cosmo, pmid, reverse = ctx.meta
        disp, vel = ctx.saved_tensors
        ptcl = Particles(conf, pmid, disp, vel=vel)
        zeros = torch.zeros_like(disp)
        cot = Particles(conf, pmid, disp_cot if disp_cot is not None else zeros,
                        vel=vel_cot if vel_cot is not None else zeros)
        _, ptcl_cot, cosmo_cot = nbody_adj(ptcl, cot, None, cosmo, conf, reverse=reverse)
        leaf_cots = tuple(cosmo_cot[n] for n in _COSMO_LEAVES if getattr(cosmo, n) is not None)
        return (ptcl_cot.disp, ptcl_cot.vel, None, None, None, None) + leaf_cots


def nbody(ptcl, obsvbl, cosmo, conf, reverse=False):
    """N-body time integration (``nbody.py:215-223``); differentiable w.r.t. ``ptcl.disp``,
    ``ptcl.vel`` and the cosmology through the reverse-time adjoint."""
    leaves = tuple(getattr(cosmo, n) for n in _COSMO_LEAVES if getattr(cosmo, n) is not None)
    disp, vel, acc = _Nbody.apply(ptcl.disp, ptcl.vel, conf, cosmo, ptcl.pmid, reverse, *leaves)
    return Particles(conf, ptcl.pmid, disp, vel=vel, acc=acc, attr=ptcl.attr), obsvbl
