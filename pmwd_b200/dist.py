"""Slab-decomposed (multi-GPU) particle-mesh path: one process per GPU, torch.distributed
(NCCL over NVLink / NVSwitch; gloo on CPU for the host-logic tests).

The reference is single-device (docs/papers/adjoint/adjoint.tex:1743-1744); this is new design
(SURVEY.md 8e):

  * mesh: x-slabs, ``Mx / P`` planes per rank; spectra after the distributed FFT are y-slabs
    ``[Mx][My / P][Mz/2+1]`` (no transpose back: the k-space kernel runs on that layout);
  * particles: the inputs and outputs of every entry point are Lagrangian x-slabs = contiguous ranges of the
    reference's C-ordered arrays (outputs concatenate to the reference order); INSIDE the integrator every particle
    lives on the rank that owns its current base plane and is re-assigned at every storage re-sort (Eulerian
    ownership, ``migrate.py`` + ``nbody._Store``; ``PMWD_MIGRATE=0`` keeps Lagrangian ownership);
  * per force: halo width (one fused pass + all-reduce) -> tiled deposit into slab + halo planes (RED kernel when
    the halo width changed since the last re-sort) -> neighbour reduce-add of halos -> local 2-D R2C over (y, z)
    -> transpose straight into the peers' symmetric-memory buffers on the copy engines -> fused x-pass on the
    y-slab layout -> 3 x (transpose, local 2-D C2R) -> neighbour halo copy -> 3-mesh gather (+ kick / drift);
    NCCL all-to-all transposes and cuFFT 1-D passes remain as fallbacks;
  * the adjoint uses the same plumbing; its float64 dot products are all-reduced once.

A slab is described to the kernels with the reference's own ``(offset, mesh shape)`` enmesh
semantics (pmwd/pm_util.py:119-141), so the CIC kernels are the single-GPU ones.
"""
import ctypes as C
import math
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .boltzmann import growth, linear_power
from .cosmology import E2
from .particles import Particles
from .scatter import make_desc


class _Timers:
    """Optional per-phase CUDA-event timers (enabled by bench / tools; off by default)."""

    def __init__(self):
        self.on = False
        self.recs = []

    class _Ctx:
        def __init__(self, owner, name):
            self.o, self.name = owner, name

        def __enter__(self):
            if self.o.on:
                self.a = torch.cuda.Event(enable_timing=True)
                self.b = torch.cuda.Event(enable_timing=True)
                self.a.record()

        def __exit__(self, *exc):
            if self.o.on:
                self.b.record()
                self.o.recs.append((self.name, self.a, self.b))

    def __call__(self, name):
        return self._Ctx(self, name)

    def read(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.recs:
            t = out.setdefault(name, [0.0, 0])
            t[0] += a.elapsed_time(b)
            t[1] += 1
        self.recs = []
        return out


TIMERS = _Timers()


class _Trace:
    """Optional event trace of one pipelined force (PMWD_TRACE=1): (label, stream) timestamps relative to
    the first mark, printed by rank 0 -- the poor man's timeline (no nsys in this image)."""

    def __init__(self):
        self.marks = []

    @property
    def on(self):
        return os.environ.get('PMWD_TRACE') == '1'

    def mark(self, label, stream=None):
        if not self.on:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else torch.cuda.current_stream())
        self.marks.append((label, ev))

    def dump(self, rank):
        if not self.marks:
            return
        torch.cuda.synchronize()
        t0 = self.marks[0][1]
        if rank == 0:
            print('   '.join(f'{lab}@{t0.elapsed_time(ev):.2f}' for lab, ev in self.marks), flush=True)
        self.marks = []


TRACE = _Trace()


class SlabComm:
    """Geometry and collectives of one rank of the slab decomposition."""

    def __init__(self, conf, group=None):
        self.conf = conf
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        P = self.size
        if conf.dim != 3:
            raise ValueError('the slab decomposition is 3-D')
        Mx, My, Mz = conf.mesh_shape
        nx, ny, nz = conf.ptcl_grid_shape
        if Mx % P or My % P or nx % P or ny % P:
            raise ValueError(f'mesh {conf.mesh_shape} / particle grid {conf.ptcl_grid_shape} '
                             f'not divisible by {P} ranks along x and y')
        self.mx, self.my = Mx // P, My // P
        self.x0, self.y0 = self.rank * self.mx, self.rank * self.my
        self.pnx = nx // P                       # particle-grid planes per rank
        self.px0 = self.rank * self.pnx
        self.ptcl_num_local = self.pnx * ny * nz
        self.left, self.right = (self.rank - 1) % P, (self.rank + 1) % P

    # ---- symmetric-memory (NVLink peer) buffers for the P2P-store transposes ---------------
    def setup_p2p(self, dev, nslots=4):
        """Allocate ``nslots`` spectrum-sized receive buffers in torch symmetric memory and
        exchange their peer-mapped pointers.  Returns False (and the NCCL all-to-all path stays
        in use) if symmetric memory is unavailable."""
        if getattr(self, '_p2p', None) is not None:
            return bool(self._p2p)
        self._p2p = []
        if os.environ.get('PMWD_P2P', '1') == '0' or self.size == 1 or self.size > 8:
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            Mx, My, Mz = self.conf.mesh_shape
            numel = Mx * self.my * (Mz // 2 + 1)
            group = dist.group.WORLD if self.group is None else self.group
            slots = []
            for _ in range(nslots):
                t = symm.empty(numel * 2, dtype=torch.float32, device=dev)
                hdl = symm.rendezvous(t, group)
                ptrs = (C.c_uint64 * self.size)(*[int(p) for p in hdl.buffer_ptrs])
                slots.append((t, hdl, ptrs))
            self._p2p = slots
            self._p2p_numel = numel
        except Exception as e:  # noqa: BLE001
            import sys
            print(f'[pmwd_b200.dist] symmetric memory unavailable ({type(e).__name__}: {e}); '
                  f'using NCCL all-to-all transposes', file=sys.stderr)
            self._p2p = []
        return bool(self._p2p)

    def p2p_barrier(self):
        self._p2p[0][1].barrier(channel=0)

    def _p2p_view(self, slot, shape):
        t = self._p2p[slot][0]
        n = shape[0] * shape[1] * shape[2]
        return torch.view_as_complex(t[:2 * n].view(n, 2)).view(shape)

    def p2p_forward(self, real, slot):
        """Local 2-D R2C + P2P-store transpose into receive buffer ``slot`` of every peer.
        The caller brackets a batch of these with :meth:`p2p_barrier`."""
        P = self.size
        mx, My, Mz = real.shape
        with TIMERS('fft2d_r2c'):
            s = self._rfft2(real)
        nzc = s.shape[-1]
        with TIMERS('p2p_transpose'):
            self._transpose(real.device, 0, mx, My // P, nzc, s, slot)
        return self._p2p_view(slot, (P * mx, My // P, nzc))

    def p2p_inverse(self, spec, slot):
        """P2P-store transpose of a y-slab spectrum ``[Mx][my][nzc]`` into buffer ``slot`` of the
        owning peers, already in the x-slab layout ``[mx][My][nzc]`` the 2-D C2R reads."""
        P = self.size
        Mx, my, nzc = spec.shape
        with TIMERS('p2p_transpose'):
            self._transpose(spec.device, 1, Mx // P, my, nzc, spec, slot)
        return self._p2p_view(slot, (Mx // P, my * P, nzc))

    def _transpose(self, dev, mode, mx, my, nzc, src, slot):
        """Slab-FFT transpose into receive buffer ``slot`` of every peer: strided peer copies on the
        copy engines (default), or the P2P-store kernel (``PMWD_P2P_CE=0``, A/B partner)."""
        lib = _lib.lib()
        st = _lib.stream_ptr(dev)
        if os.environ.get('PMWD_P2P_CE', '1') != '0':
            _lib.check(lib.pmwd_transpose_ce(st, mode, self.size, self.rank, mx, my, nzc, _lib.ptr(src),
                                             self._p2p[slot][2], int(os.environ.get('PMWD_P2P_CE_STREAMS', '4'))),
                       'pmwd_transpose_ce')
        else:
            _lib.check(lib.pmwd_transpose_p2p(st, mode, self.size, self.rank, mx, my, nzc, _lib.ptr(src),
                                              self._p2p[slot][2]), 'pmwd_transpose_p2p')

    # ---- particles -----------------------------------------------------------------------
    def local_slice(self):
        """Range of this rank in the reference's (global, C-ordered) particle arrays."""
        lo = self.rank * self.ptcl_num_local
        return slice(lo, lo + self.ptcl_num_local)

    # ---- all-to-all transposes of the distributed FFT ------------------------------------
    def _a2a(self, send):
        send = send.contiguous()
        recv = torch.empty_like(send)
        s, r = (torch.view_as_real(send), torch.view_as_real(recv)) if send.is_complex() else (send, recv)
        dist.all_to_all_single(r, s, group=self.group)
        return recv

    def _fft_x(self, s, inverse):
        """1-D complex FFT along dim 0 of a contiguous [Mx][my][nzc] array, in place on CUDA
        (cuFFT strided-batch plan behind ``pmwd_fft_c2c_lead``; torch.fft on CPU for the gloo
        host-logic tests)."""
        if not s.is_cuda:
            return (torch.fft.ifft(s, dim=0, norm='forward') if inverse else torch.fft.fft(s, dim=0)).contiguous()
        s = s.contiguous()
        ctx = _lib.Context.get(s.device)
        with torch.cuda.device(s.device):
            _lib.check(_lib.lib().pmwd_fft_c2c_lead(ctx.handle, _lib.stream_ptr(s.device), s.shape[0],
                                                    s.shape[1] * s.shape[2], _lib.ptr(s), int(inverse)),
                       'pmwd_fft_c2c_lead')
        return s

    def _rfft2(self, real):
        """Local 2-D R2C over (y, z) of a contiguous [mx][My][Mz] slab (cuFFT plans of the C
        library on CUDA; torch.fft on CPU for the gloo tests)."""
        if not real.is_cuda:
            return torch.fft.rfft2(real)
        real = real.contiguous()
        shape = tuple(real.shape)
        ctx = _lib.Context.get(real.device).reserve(shape)
        out = torch.empty(shape[:2] + (shape[2] // 2 + 1,), dtype=torch.complex64, device=real.device)
        with torch.cuda.device(real.device):
            _lib.check(_lib.lib().pmwd_fft2d_r2c(ctx.handle, _lib.stream_ptr(real.device), _lib.shape_arr(shape),
                                                 _lib.ptr(real), _lib.ptr(out)), 'pmwd_fft2d_r2c')
        return out

    def _irfft2(self, s, My, Mz, out=None):
        """Local 2-D C2R over (y, z), unnormalised; ``s`` (contiguous) is clobbered on CUDA."""
        if not s.is_cuda:
            r = torch.fft.irfft2(s, s=(My, Mz), norm='forward')
            if out is not None:
                out.copy_(r)
                return out
            return r
        shape = (s.shape[0], My, Mz)
        ctx = _lib.Context.get(s.device).reserve(shape)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=s.device)
        assert out.is_contiguous() and s.is_contiguous()
        with torch.cuda.device(s.device):
            _lib.check(_lib.lib().pmwd_fft2d_c2r(ctx.handle, _lib.stream_ptr(s.device), _lib.shape_arr(shape),
                                                 _lib.ptr(s), _lib.ptr(out)), 'pmwd_fft2d_c2r')
        return out

    def rfft2_a2a(self, real):
        """x-slab real ``[mx][My][Mz]`` -> ``[Mx][my][Mz/2+1]`` transformed over (y, z) only and
        transposed to y-slabs (the input layout of the fused x-pass, csrc/xpass.cu)."""
        P = self.size
        mx, My, Mz = real.shape
        my = My // P
        with TIMERS('fft2d_r2c'):
            s = self._rfft2(real)                                   # local 2-D R2C over (y, z)
        nzc = s.shape[-1]
        with TIMERS('pack'):
            s = s.reshape(mx, P, my, nzc).permute(1, 0, 2, 3).contiguous()   # pack per destination
        with TIMERS('all_to_all'):
            return self._a2a(s).reshape(P * mx, my, nzc)            # blocks arrive in x order

    def a2a_irfft2(self, s, My, Mz, out=None):
        """Inverse of :meth:`rfft2_a2a` without any 1/N: ``[Mx][my][nzc]`` -> real ``[mx][My][Mz]``."""
        P = self.size
        Mx, my, nzc = s.shape
        mx = Mx // P
        with TIMERS('all_to_all'):
            s = self._a2a(s.reshape(P, mx, my, nzc))                # [p] = y-chunk p of my planes
        with TIMERS('pack'):
            s = s.permute(1, 0, 2, 3).reshape(mx, My, nzc)          # unpack (copies)
        with TIMERS('fft2d_c2r'):
            return self._irfft2(s, My, Mz, out=out)

    def a2a_start(self, s):
        """Start the transpose of :meth:`a2a_irfft2` asynchronously (NCCL stream); returns a
        token for :meth:`a2a_finish_irfft2`.  Lets the 2-D C2R of one force component overlap
        the all-to-all of the next."""
        P = self.size
        Mx, my, nzc = s.shape
        send = s.reshape(P, Mx // P, my, nzc).contiguous()
        recv = torch.empty_like(send)
        work = dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send),
                                      group=self.group, async_op=True)
        return send, recv, work

    def a2a_finish_irfft2(self, token, My, Mz, out=None):
        send, recv, work = token
        with TIMERS('all_to_all_wait'):
            work.wait()
        P, mx, my, nzc = recv.shape
        with TIMERS('pack'):
            s = recv.permute(1, 0, 2, 3).reshape(mx, My, nzc)       # unpack (copies)
        with TIMERS('fft2d_c2r'):
            return self._irfft2(s, My, Mz, out=out)

    def rfftn(self, real, shape=None):
        """x-slab real ``[mx][My][Mz]`` -> y-slab spectrum ``[Mx][my][Mz/2+1]`` (unnormalised,
        = numpy rfftn of the global field, pmwd/pm_util.py:281)."""
        s = self.rfft2_a2a(real)
        with TIMERS('fft1d_x'):
            return self._fft_x(s, inverse=False)                    # 1-D C2C over x

    def irfftn(self, spec, My, Mz, out=None):
        """Inverse of :meth:`rfftn` WITHOUT the 1/N (callers fold it into their scale).
        ``spec`` is clobbered on CUDA (in-place x-pass)."""
        with TIMERS('fft1d_x'):
            s = self._fft_x(spec, inverse=True)                     # unnormalised inverse over x (in place)
        return self.a2a_irfft2(s, My, Mz, out=out)

    # ---- halos ---------------------------------------------------------------------------
    def _exchange(self, to_left, to_right):
        """Send one tensor to each x-neighbour, receive theirs: (from_left, from_right)."""
        from_left, from_right = torch.empty_like(to_right), torch.empty_like(to_left)
        if self.size == 1:
            from_left.copy_(to_right); from_right.copy_(to_left)
            return from_left, from_right
        ops = [dist.P2POp(dist.isend, to_right, self.right, self.group),
               dist.P2POp(dist.isend, to_left, self.left, self.group),
               dist.P2POp(dist.irecv, from_left, self.left, self.group),
               dist.P2POp(dist.irecv, from_right, self.right, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return from_left, from_right

    def halo_reduce(self, ext, h):
        """``ext[..., mx + 2h, My, Mz]`` deposited with halos -> adds the halo planes into the
        owning neighbours; returns the owned slab view ``ext[..., h:h+mx]``."""
        mx = self.mx
        lo = ext[..., :h, :, :].contiguous()           # planes x0-h .. x0-1   -> left neighbour
        hi = ext[..., h + mx:, :, :].contiguous()      # planes x0+mx ..       -> right neighbour
        from_left, from_right = self._exchange(lo, hi)
        ext[..., h:2 * h, :, :] += from_left            # their upper halo = my first planes
        ext[..., mx:mx + h, :, :] += from_right         # their lower halo = my last planes
        return ext[..., h:h + mx, :, :]

    def halo_fill(self, ext, h):
        """Owned planes ``ext[..., h:h+mx]`` are valid; fetch the halos from the neighbours."""
        mx = self.mx
        first = ext[..., h:2 * h, :, :].contiguous()    # -> left neighbour's upper halo
        last = ext[..., mx:mx + h, :, :].contiguous()   # -> right neighbour's lower halo
        from_left, from_right = self._exchange(first, last)
        ext[..., :h, :, :] = from_left
        ext[..., h + mx:, :, :] = from_right
        return ext

    def allreduce_max(self, x):
        t = x.detach().reshape(1).clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t)

    def allreduce_sum_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


# ------------------------------------------------------------------------------------------
class SlabForce:
    """``gravity`` (pmwd/gravity.py:47-72) and ``force_adj`` (pmwd/nbody.py:108-118) on slabs."""

    def __init__(self, conf, comm):
        self.conf, self.comm = conf, comm
        self.h_alloc = 0
        self._ext1 = self._ext3 = None
        self._F = None           # force meshes with halos kept for the adjoint gather
        self._h = None

    def _side_stream(self, dev):
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    def _halo(self, pmid, disp):
        """Halo planes needed on either side of the slab: how far the base plane of any local particle
        (``pmid_x + floor(disp_x / cell)``, float32 arithmetic of ``pmwd/pm_util.py:129-136``) lies
        outside the owned planes, + the stencil's second plane.  With Lagrangian ownership this grows
        with the largest displacement; with Eulerian ownership (``nbody._Store.migrator``) it stays at
        the drift accumulated since the last migration."""
        conf, comm = self.conf, self.comm
        Mx, mx = conf.mesh_shape[0], comm.mx
        if disp.is_cuda and pmid.dtype == torch.int16:
            m = torch.empty(1, dtype=torch.int32, device=disp.device)
            with torch.cuda.device(disp.device):
                _lib.check(_lib.lib().pmwd_slab_owner(_lib.stream_ptr(disp.device), pmid.shape[0], _lib.ptr(pmid),
                                                      _lib.ptr(disp), float(conf.cell_size), Mx, comm.size, comm.x0, mx,
                                                      None, _lib.ptr(m)), 'pmwd_slab_owner')
        else:
            cell32 = float(np.float32(conf.cell_size))
            plane = pmid[:, 0].to(torch.int32) + torch.floor(disp[:, 0] / cell32).to(torch.int32)
            d = torch.remainder(plane - comm.x0, Mx)                 # 0 .. Mx-1; owned if d < mx
            right = d + (2 - mx)                                     # planes needed above the slab
            need = torch.where(d < mx, right.clamp(min=0), torch.minimum(right, Mx - d))
            m = need.max() if need.numel() else torch.zeros((), dtype=torch.int32, device=disp.device)
        h = max(int(comm.allreduce_max(m.to(torch.float32))), 1)
        if h > comm.mx:
            raise RuntimeError(f'halo of {h} planes exceeds the slab width {comm.mx}: use fewer ranks')
        if h > self.h_alloc:
            self.h_alloc = min(comm.mx, (h + 7) // 8 * 8)
            self._ext1 = self._ext3 = None
        return self.h_alloc

    def _desc(self, pmid, h):
        conf, comm = self.conf, self.comm
        _, My, Mz = conf.mesh_shape
        # whole planes of the float32 cell size enmesh divides by (pm_util.py:120: divmod(b12, a1)
        # with a1 = float32(cell_size)), so that the remainder is exactly zero for ANY cell size,
        # also one that float32 cannot represent (1/3, 0.05, ...): the fast-path kernels need it
        off = ((comm.x0 - h) * float(np.float32(conf.cell_size)), 0.0, 0.0)
        return make_desc(conf, pmid, (comm.mx + 2 * h, My, Mz), 1, off, None)

    def _buffers(self, dev, h):
        conf, comm = self.conf, self.comm
        _, My, Mz = conf.mesh_shape
        shape = (comm.mx + 2 * h, My, Mz)
        if self._ext1 is None or self._ext1.shape != shape:
            self._ext1 = torch.empty(shape, dtype=conf.float_dtype, device=dev)
            self._ext3 = torch.empty((3,) + shape, dtype=conf.float_dtype, device=dev)
        return self._ext1, self._ext3

    def _mesh_forces(self, pmid, disp, Om, h, sweep=None):
        """Particles -> the three force meshes with halos, ``[3][mx+2h][My][Mz]``.  ``sweep``: the store's
        ``sweep.SweepState``; if its table was built for this very slab + halo the deposit goes through the
        tiled kernels (csrc/scatter_sweep.cu), else through the per-particle RED kernel."""
        conf, comm = self.conf, self.comm
        lib = _lib.lib()
        dev = disp.device
        Mx, My, Mz = conf.mesh_shape
        ext1, ext3 = self._buffers(dev, h)
        desc = self._desc(pmid, h)
        st = _lib.stream_ptr(dev)
        val = float(np.float32(conf.mesh_size / conf.ptcl_num))       # scatter.py:37-39
        self._swept = sweep is not None and sweep.usable(desc)
        with TIMERS('scatter'):
            if self._swept:
                _lib.check(lib.pmwd_scatter_sweep(st, C.byref(desc), sweep.arg(), _lib.ptr(pmid), _lib.ptr(disp), None,
                                                  val, 1, _lib.ptr(ext1), None, None), 'pmwd_scatter_sweep')
            else:
                ext1.zero_()
                _lib.check(lib.pmwd_scatter_soa(st, C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp), None, val, 1,
                                                _lib.ptr(ext1), None, None), 'pmwd_scatter_soa')
        with TIMERS('halo'):
            rho = comm.halo_reduce(ext1, h)
        fused = bool(lib.pmwd_xpass_supported(Mx)) and os.environ.get('PMWD_XPASS', '1') != '0'
        p2p = fused and disp.is_cuda and comm.setup_p2p(dev)
        npipe = int(os.environ.get('PMWD_PIPE', '1'))
        if p2p and npipe > 1 and os.environ.get('PMWD_P2P_CE', '1') != '0' and comm.mx % npipe == 0 \
                and comm.my % npipe == 0 and ((comm.my // npipe) * (Mz // 2 + 1)) % 2 == 0:
            self._spectral_pipelined(rho, ext3, h, Om, npipe)
            with TIMERS('halo'):
                comm.halo_fill(ext3, h)
            return desc, ext3, val
        if p2p:
            comm.p2p_barrier()                      # peers are done reading their receive buffers
            spec = comm.p2p_forward(rho, 0)
            comm.p2p_barrier()                      # every rank's rows have landed
        else:
            spec = comm.rfft2_a2a(rho) if fused else comm.rfftn(rho)
        g = [torch.empty_like(spec) for _ in range(3)]
        scale = float(np.float32(1.5 * Om / conf.mesh_size))          # 1.5 Omega_m and irfftn's 1/N
        arr = (C.c_void_p * 3)(*[t.data_ptr() for t in g])
        with TIMERS('kspace'):
            fn = lib.pmwd_xpass_force if fused else lib.pmwd_kspace_force_slab
            _lib.check(fn(st, _lib.shape_arr(conf.mesh_shape), comm.y0, comm.my, float(conf.cell_size), scale,
                          _lib.ptr(spec), arr), 'pmwd_xpass_force / pmwd_kspace_force_slab')
        del spec
        if p2p:
            # transposes on a side stream, one cross-rank barrier per component; the 2-D C2R of
            # component i (main stream) overlaps the NVLink stores of component i+1
            main = torch.cuda.current_stream(dev)
            side = self._side_stream(dev)
            side.wait_stream(main)
            landed, recv = [], []
            with torch.cuda.stream(side):
                comm.p2p_barrier()                  # peers are done with their receive buffers
                for i in range(3):
                    recv.append(comm.p2p_inverse(g[i], 1 + i))
                    comm.p2p_barrier()              # component i has landed everywhere
                    ev = torch.cuda.Event()
                    ev.record(side)
                    landed.append(ev)
            for i in range(3):
                main.wait_event(landed[i])
                with TIMERS('fft2d_c2r'):
                    comm._irfft2(recv[i], My, Mz, out=ext3[i, h:h + comm.mx])
            for t in g:
                t.record_stream(side)
            g = None
        elif fused and dist.get_backend(comm.group) == 'nccl':
            tokens = [comm.a2a_start(g[i]) for i in range(3)]      # comm of i+1 overlaps C2R of i
            g = None
            for i in range(3):
                comm.a2a_finish_irfft2(tokens[i], My, Mz, out=ext3[i, h:h + comm.mx])
                tokens[i] = None
        else:
            for i in range(3):
                if fused:
                    comm.a2a_irfft2(g[i], My, Mz, out=ext3[i, h:h + comm.mx])
                else:
                    comm.irfftn(g[i], My, Mz, out=ext3[i, h:h + comm.mx])
                g[i] = None
        with TIMERS('halo'):
            comm.halo_fill(ext3, h)
        return desc, ext3, val

    def _spectral_pipelined(self, rho, ext3, h, Om, npipe):
        """rho ``[mx][My][Mz]`` -> the three force meshes in ``ext3[:, h:h+mx]``, with the slab-FFT
        transposes (copy engines, side stream) hidden under the transforms: the 2-D R2C runs on
        ``npipe`` chunks of x planes, each chunk's rows leaving for the peers while the next chunk
        is transformed; the x-pass runs on ``npipe`` y-parts of the y-slab, each part's three
        spectra leaving while the next part is computed; component i's 2-D C2R starts as soon as
        its last part has landed everywhere."""
        conf, comm = self.conf, self.comm
        lib = _lib.lib()
        dev = rho.device
        Mx, My, Mz = conf.mesh_shape
        P, rank, mx, my = comm.size, comm.rank, comm.mx, comm.my
        nzc = Mz // 2 + 1
        myh, mxc = my // npipe, mx // npipe
        run_h = myh * nzc * 8                       # bytes of one (x plane, y-part) run
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev)
        ns = int(os.environ.get('PMWD_P2P_CE_STREAMS', '4'))
        slot_ptrs = [comm._p2p[k][2] for k in range(4)]
        scale = float(np.float32(1.5 * Om / conf.mesh_size))

        def copy(src_ptr, width, height, src_peer_stride, spitch, slot, dst_off, dpitch):
            _lib.check(lib.pmwd_peer_copy2d(_lib.stream_ptr(dev), P, rank, width, height, C.c_void_p(src_ptr),
                                            src_peer_stride, spitch, slot_ptrs[slot], dst_off, dpitch, ns),
                       'pmwd_peer_copy2d')

        TRACE.mark('start', main)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            comm.p2p_barrier()                      # peers are done with all their receive buffers
            TRACE.mark('bar0', side)
        # ---- forward: R2C per x-chunk (main), rows to the peers' y-part buffers (side)
        keep = []
        for c in range(npipe):
            with TIMERS('fft2d_r2c'):
                s_c = comm._rfft2(rho[c * mxc:(c + 1) * mxc])
            ev = torch.cuda.Event(); ev.record(main)
            TRACE.mark(f'r2c{c}', main)
            keep.append(s_c)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                for hh in range(npipe):
                    copy(s_c.data_ptr() + hh * run_h, run_h, mxc, my * nzc * 8, My * nzc * 8, 0,
                         hh * Mx * run_h + (rank * mx + c * mxc) * run_h, run_h)
                TRACE.mark(f'T{c}', side)
            s_c.record_stream(side)
        with torch.cuda.stream(side):
            comm.p2p_barrier()                      # every rank's rows have landed
            ev_fwd = torch.cuda.Event(); ev_fwd.record(side)
            TRACE.mark('bar1', side)
        main.wait_event(ev_fwd)
        keep = None
        # ---- x-pass per y-part (main); its three spectra to the owners of the x planes (side)
        recv0 = comm._p2p[0][0]
        g = [[torch.empty((Mx, myh, nzc), dtype=torch.complex64, device=dev) for _ in range(3)] for _ in range(npipe)]
        evx = []
        for hh in range(npipe):
            n = Mx * myh * nzc
            spec_h = torch.view_as_complex(recv0[2 * hh * n:2 * (hh + 1) * n].view(n, 2)).view(Mx, myh, nzc)
            arr = (C.c_void_p * 3)(*[t.data_ptr() for t in g[hh]])
            with TIMERS('kspace'):
                _lib.check(lib.pmwd_xpass_force(_lib.stream_ptr(dev), _lib.shape_arr(conf.mesh_shape), comm.y0 + hh * myh,
                                                myh, float(conf.cell_size), scale, _lib.ptr(spec_h), arr),
                           'pmwd_xpass_force')
            ev = torch.cuda.Event(); ev.record(main)
            TRACE.mark(f'xp{hh}', main)
            evx.append(ev)
        landed = []
        with torch.cuda.stream(side):
            # component by component, so that the first 2-D C2R can start as early as possible
            for i in range(3):
                for hh in range(npipe):
                    if i == 0:
                        side.wait_event(evx[hh])
                    copy(g[hh][i].data_ptr(), run_h, mx, mx * run_h, run_h, 1 + i,
                         (rank * my + hh * myh) * nzc * 8, My * nzc * 8)
                comm.p2p_barrier()                  # component i has landed everywhere
                ev = torch.cuda.Event(); ev.record(side)
                TRACE.mark(f'L{i}', side)
                landed.append(ev)
        for part in g:
            for t in part:
                t.record_stream(side)
        for i in range(3):
            main.wait_event(landed[i])
            with TIMERS('fft2d_c2r'):
                comm._irfft2(comm._p2p_view(1 + i, (mx, My, nzc)), My, Mz, out=ext3[i, h:h + mx])
            TRACE.mark(f'c2r{i}', main)
        TRACE.dump(rank)

    def _spectral_adj_pipelined(self, Vs, rc, h, Om, npipe):
        """The mesh part of ``force_adj``: the three cotangent meshes ``Vs[i]`` ``[mx][My][Mz]`` -> rho_cot in
        ``rc[h:h+mx]``.  Component i's rows leave for the peers (copy engines, side stream) while component
        i+1 is transformed; the x-pass runs on ``npipe`` y-parts whose output leaves while the next part is
        computed."""
        conf, comm = self.conf, self.comm
        lib = _lib.lib()
        dev = rc.device
        Mx, My, Mz = conf.mesh_shape
        P, rank, mx, my = comm.size, comm.rank, comm.mx, comm.my
        nzc = Mz // 2 + 1
        myh = my // npipe
        run_h = myh * nzc * 8
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev)
        ns = int(os.environ.get('PMWD_P2P_CE_STREAMS', '4'))
        slot_ptrs = [comm._p2p[k][2] for k in range(4)]
        scale = float(np.float32(1.5 * Om / conf.mesh_size))

        def copy(src_ptr, width, height, src_peer_stride, spitch, slot, dst_off, dpitch):
            _lib.check(lib.pmwd_peer_copy2d(_lib.stream_ptr(dev), P, rank, width, height, C.c_void_p(src_ptr),
                                            src_peer_stride, spitch, slot_ptrs[slot], dst_off, dpitch, ns),
                       'pmwd_peer_copy2d')

        side.wait_stream(main)
        with torch.cuda.stream(side):
            comm.p2p_barrier()
        for i in range(3):
            with TIMERS('fft2d_r2c'):
                s_i = comm._rfft2(Vs[i])
            ev = torch.cuda.Event(); ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                for hh in range(npipe):
                    copy(s_i.data_ptr() + hh * run_h, run_h, mx, my * nzc * 8, My * nzc * 8, i,
                         hh * Mx * run_h + rank * mx * run_h, run_h)
            s_i.record_stream(side)
            del s_i
        with torch.cuda.stream(side):
            comm.p2p_barrier()
            ev_fwd = torch.cuda.Event(); ev_fwd.record(side)
        main.wait_event(ev_fwd)
        n = Mx * myh * nzc
        outs, evx = [], []
        for hh in range(npipe):
            parts = [torch.view_as_complex(comm._p2p[i][0][2 * hh * n:2 * (hh + 1) * n].view(n, 2)).view(Mx, myh, nzc)
                     for i in range(3)]
            out_h = torch.empty((Mx, myh, nzc), dtype=torch.complex64, device=dev)
            arr = (C.c_void_p * 3)(*[t.data_ptr() for t in parts])
            with TIMERS('kspace'):
                _lib.check(lib.pmwd_xpass_force_adj(_lib.stream_ptr(dev), _lib.shape_arr(conf.mesh_shape),
                                                    comm.y0 + hh * myh, myh, float(conf.cell_size), scale, arr,
                                                    _lib.ptr(out_h)), 'pmwd_xpass_force_adj')
            ev = torch.cuda.Event(); ev.record(main)
            outs.append(out_h); evx.append(ev)
        with torch.cuda.stream(side):
            for hh in range(npipe):
                side.wait_event(evx[hh])
                copy(outs[hh].data_ptr(), run_h, mx, mx * run_h, run_h, 3, (rank * my + hh * myh) * nzc * 8,
                     My * nzc * 8)
            comm.p2p_barrier()
            landed = torch.cuda.Event(); landed.record(side)
        for t in outs:
            t.record_stream(side)
        main.wait_event(landed)
        with TIMERS('fft2d_c2r'):
            comm._irfft2(comm._p2p_view(3, (mx, My, nzc)), My, Mz, out=rc[h:h + mx])

    def force(self, pmid, disp, Om, acc, kick_vel=None, kick_factor=0.0, next_kd=None, sweep=None):
        """``next_kd = (K1_next, D_next)``: also apply the next step's leading half-kick and
        drift in the gather pass (pipelined KDK, cf. ``pmwd_force_kdk``)."""
        with TIMERS('halo_width'):
            h = self._halo(pmid, disp)
        desc, F, _ = self._mesh_forces(pmid, disp, Om, h, sweep)
        with TIMERS('gather'):
            st = _lib.stream_ptr(disp.device)
            if next_kd is not None:
                _lib.check(_lib.lib().pmwd_gather3_kdk(
                    st, C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp), _lib.ptr(F[0]), _lib.ptr(F[1]),
                    _lib.ptr(F[2]), _lib.ptr(acc), _lib.ptr(kick_vel), float(kick_factor), float(next_kd[0]),
                    float(next_kd[1])), 'pmwd_gather3_kdk')
            else:
                _lib.check(_lib.lib().pmwd_gather3(
                    st, C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp), _lib.ptr(F[0]), _lib.ptr(F[1]),
                    _lib.ptr(F[2]), _lib.ptr(acc), _lib.ptr(kick_vel), float(kick_factor)), 'pmwd_gather3')

    def force_adj(self, pmid, disp, Om, pi, acc, alpha, sweep=None):
        conf, comm = self.conf, self.comm
        lib = _lib.lib()
        dev = disp.device
        Mx, My, Mz = conf.mesh_shape
        h = self._halo(pmid, disp)
        desc, F, val = self._mesh_forces(pmid, disp, Om, h, sweep)
        st = _lib.stream_ptr(dev)
        _lib.check(lib.pmwd_gather3(st, C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp), _lib.ptr(F[0]),
                                    _lib.ptr(F[1]), _lib.ptr(F[2]), _lib.ptr(acc), None, 0.0), 'pmwd_gather3')
        # V_i = scatter(pi_i) (gather.py:113), halos reduced
        if self._swept:
            # three tiled deposits (they overwrite V; the straggler list of the density deposit is re-recorded
            # by the first one and reused by the other two)
            V = torch.empty_like(F)
            _lib.check(lib.pmwd_scatter_sweep(st, C.byref(desc), sweep.arg(), _lib.ptr(pmid), _lib.ptr(disp),
                                              _lib.ptr(pi), 0.0, 3, _lib.ptr(V[0]), _lib.ptr(V[1]), _lib.ptr(V[2])),
                       'pmwd_scatter_sweep')
        else:
            V = torch.zeros_like(F)
            _lib.check(lib.pmwd_scatter_soa(st, C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp), _lib.ptr(pi), 0.0, 3,
                                            _lib.ptr(V[0]), _lib.ptr(V[1]), _lib.ptr(V[2])), 'pmwd_scatter_soa')
        Vs = comm.halo_reduce(V, h)
        fused = bool(lib.pmwd_xpass_supported(Mx)) and os.environ.get('PMWD_XPASS', '1') != '0'
        p2p = fused and comm.setup_p2p(dev)
        npipe = int(os.environ.get('PMWD_PIPE', '1'))
        if p2p and npipe > 1 and os.environ.get('PMWD_P2P_CE', '1') != '0' and comm.my % npipe == 0 \
                and ((comm.my // npipe) * (Mz // 2 + 1)) % 2 == 0:
            rc = torch.empty_like(F[0])
            self._spectral_adj_pipelined(Vs, rc, h, Om, npipe)
            del V, Vs
            with TIMERS('halo'):
                comm.halo_fill(rc, h)
            _lib.check(lib.pmwd_force_adj_gather(st, C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp), _lib.ptr(F[0]),
                                                 _lib.ptr(F[1]), _lib.ptr(F[2]), _lib.ptr(rc), _lib.ptr(pi), val,
                                                 _lib.ptr(alpha)), 'pmwd_force_adj_gather')
            return
        if p2p:
            comm.p2p_barrier()
            S = [comm.p2p_forward(Vs[i].contiguous(), i) for i in range(3)]
            comm.p2p_barrier()
        else:
            S = [(comm.rfft2_a2a if fused else comm.rfftn)(Vs[i].contiguous()) for i in range(3)]
        del V, Vs
        out = torch.empty_like(S[0])
        scale = float(np.float32(1.5 * Om / conf.mesh_size))
        arr = (C.c_void_p * 3)(*[t.data_ptr() for t in S])
        fn = lib.pmwd_xpass_force_adj if fused else lib.pmwd_kspace_force_adj_slab
        _lib.check(fn(st, _lib.shape_arr(conf.mesh_shape), comm.y0, comm.my, float(conf.cell_size), scale, arr,
                      _lib.ptr(out)), 'pmwd_xpass_force_adj / pmwd_kspace_force_adj_slab')
        del S
        rc = torch.empty_like(F[0])
        if p2p:
            r = comm.p2p_inverse(out, 3)
            comm.p2p_barrier()
            comm._irfft2(r, My, Mz, out=rc[h:h + comm.mx])
        elif fused:
            comm.a2a_irfft2(out, My, Mz, out=rc[h:h + comm.mx])
        else:
            comm.irfftn(out, My, Mz, out=rc[h:h + comm.mx])
        comm.halo_fill(rc, h)
        _lib.check(lib.pmwd_force_adj_gather(st, C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp), _lib.ptr(F[0]),
                                             _lib.ptr(F[1]), _lib.ptr(F[2]), _lib.ptr(rc), _lib.ptr(pi), val,
                                             _lib.ptr(alpha)), 'pmwd_force_adj_gather')


# ------------------------------------------------------------------------------------------
def _kvec_T(conf, comm, shape, spacing, dev):
    """Wavevectors of the transposed slab layout [n0][n1/P][n2/2+1] (pm_util.py:159-199)."""
    P, r = comm.size, comm.rank
    period = 2 * math.pi / spacing
    k0 = torch.from_numpy(np.fft.fftfreq(shape[0]) * period).to(conf.float_dtype).to(dev).reshape(-1, 1, 1)
    n1l = shape[1] // P
    k1 = np.fft.fftfreq(shape[1])[r * n1l:(r + 1) * n1l] * period
    k1 = torch.from_numpy(k1).to(conf.float_dtype).to(dev).reshape(1, -1, 1)
    k2 = torch.from_numpy(np.fft.rfftfreq(shape[2]) * period).to(conf.float_dtype).to(dev).reshape(1, 1, -1)
    return [k0, k1, k2]


def white_noise_slab(seed, conf, comm, dev, exact=True):
    """This rank's x-slab of the real white-noise field.  ``exact``: slice of the single-GPU
    stream (numpy default_rng(seed) over the whole grid; for parity tests); otherwise an
    independent per-rank stream (large runs)."""
    nx, ny, nz = conf.ptcl_grid_shape
    if exact:
        full = np.random.default_rng(seed).standard_normal(conf.ptcl_grid_shape, dtype=np.float32)
        return torch.from_numpy(full[comm.px0:comm.px0 + comm.pnx].copy()).to(dev)
    g = torch.Generator(device=dev).manual_seed(seed * 1000003 + comm.rank)
    return torch.randn((comm.pnx, ny, nz), device=dev, generator=g, dtype=conf.float_dtype)


def lpt_slab(white, cosmo, conf, comm):
    """``linear_modes`` + ``lpt`` (pmwd/modes.py:67-84, pmwd/lpt.py:136-212) on particle-grid
    slabs from a real white-noise slab; elementwise k-space work in torch (run once, off the
    hot path).  Returns this rank's ``Particles`` (Lagrangian x-slab)."""
    dev = white.device
    fdt = conf.float_dtype
    nx, ny, nz = conf.ptcl_grid_shape
    N = conf.ptcl_num
    kvec = _kvec_T(conf, comm, conf.ptcl_grid_shape, conf.ptcl_spacing, dev)
    k2 = kvec[0] ** 2 + kvec[1] ** 2 + kvec[2] ** 2
    kk = torch.sqrt(k2)
    Plin = linear_power(kk, None, cosmo, conf)
    modes = comm.rfftn(white) / math.sqrt(N)                        # norm='ortho'
    modes = modes * torch.sqrt(Plin * conf.box_vol).to(fdt)
    modes = modes / conf.ptcl_cell_vol
    inv = torch.where(k2 != 0, -1 / torch.where(k2 != 0, k2, torch.ones_like(k2)), torch.zeros_like(k2))
    nyq = math.pi / conf.ptcl_spacing
    eps = nyq * torch.finfo(fdt).eps
    km = [torch.where((k.abs() - nyq).abs() <= eps, torch.zeros_like(k), k) for k in kvec]

    def inv_fft(s):
        return comm.irfftn(s, ny, nz) / N

    pot = [modes * inv]
    if conf.lpt_order > 1:
        def strain(i, j):
            m = -kvec[i] * kvec[j] if i == j else -km[i] * km[j]
            return inv_fft(m * pot[0])
        d = [strain(i, i) for i in range(3)]
        L = d[0] * d[2] + d[0] * d[1] + d[1] * d[2]
        for i in range(2):
            for j in range(i + 1, 3):
                s = strain(i, j)
                L = L - s * s
        pot.append(comm.rfftn(L) * inv)
    if conf.lpt_order > 2:
        raise NotImplementedError('TODO')

    # gen_grid for this slab (particles.py:109-144); integer mesh ratio assumed exact here
    full = Particles.gen_grid(conf.replace(ptcl_grid_shape=(1, ny, nz),
                                           mesh_shape=(conf.mesh_shape[0] // nx,) + conf.mesh_shape[1:]),
                              device=dev)
    ratio = conf.mesh_shape[0] / nx
    ix = torch.arange(comm.px0, comm.px0 + comm.pnx, device=dev)
    pm_x = torch.round(ix * ratio).to(conf.pmid_dtype)
    dx = ((ix * conf.mesh_shape[0] - pm_x.to(torch.int64) * nx) * (conf.cell_size / nx)).to(fdt)
    pmid = full.pmid.reshape(1, ny * nz, 3).repeat(comm.pnx, 1, 1)
    pmid[:, :, 0] = pm_x[:, None]
    disp = full.disp.reshape(1, ny * nz, 3).repeat(comm.pnx, 1, 1)
    disp[:, :, 0] = dx[:, None]
    pmid = pmid.reshape(-1, 3).contiguous()
    disp = disp.reshape(-1, 3).contiguous()
    vel = torch.zeros_like(disp)

    a = conf.a_start
    for order in range(1, 1 + conf.lpt_order):
        D = growth(a, cosmo, conf, order=order)
        a2HDp = a ** 2 * torch.sqrt(E2(a, cosmo)) * growth(a, cosmo, conf, order=order, deriv=1)
        D = float(np.float32(float(D)))
        a2HDp = float(np.float32(float(a2HDp)))
        for i in range(3):
            grad = inv_fft(-1j * km[i] * pot[order - 1]).to(fdt).reshape(-1)
            disp[:, i] += D * grad
            vel[:, i] += a2HDp * grad
    return Particles(conf, pmid, disp, vel=vel)


# ------------------------------------------------------------------------------------------
def nbody_slab(ptcl, cosmo, conf, comm, reverse=False, force=None):
    """``nbody`` (pmwd/nbody.py:215-223) on this rank's Lagrangian slab of particles."""
    from .nbody import _Store, _owned, drift_factor, kick_factor, _f32, _kick_drift
    force = force or SlabForce(conf, comm)
    a_nbody = conf.a_nbody.tolist()
    if reverse:
        a_nbody = a_nbody[::-1]
    Om = float(cosmo.Omega_m)
    with torch.no_grad():
        p = _owned(ptcl, conf)
        store = _Store(conf, dict(pmid=p.pmid.clone(), disp=p.disp, vel=p.vel, acc=p.acc))
        stepper = SlabStepper(store, a_nbody, cosmo, conf, comm, force)
        stepper.init()
        for _ in range(stepper.nsteps):
            stepper.step()
        disp, vel, acc = store.lagrangian('disp', 'vel', 'acc')
    return Particles(conf, ptcl.pmid, disp, vel=vel, acc=acc)


class SlabStepper:
    """Pipelined KDK stepping on one rank's slab (cf. ``nbody._Stepper``): one force per step
    with the kick/drift of both adjacent half-steps applied in its gather pass."""

    def __init__(self, store, a_list, cosmo, conf, comm, force):
        from .nbody import _Stepper
        self.inner = _Stepper(store, a_list, cosmo, conf)     # factor cache only
        self.store, self.cosmo, self.conf, self.comm, self.force = store, cosmo, conf, comm, force
        if tuple(tuple(x) for x in conf.symp_splits) != ((0, 0.5), (1, 0.5)):
            raise NotImplementedError('the slab integrator implements the default KDK splitting')
        self.i = 0
        self.pre = False
        store.desc_fn = lambda pmid: force._desc(pmid, max(force.h_alloc, 1))
        if os.environ.get('PMWD_MIGRATE', '1') != '0' and comm.size > 1:
            store.migrator = comm              # Eulerian ownership: re-assigned at every storage re-sort
            # ... which keeps the halos narrow and stable: the tiled deposit's table stays valid between re-sorts
            store.slab_sweep = os.environ.get('PMWD_SLAB_SWEEP', '1') != '0'
            store.timers = TIMERS

    @property
    def nsteps(self):
        return self.inner.nsteps

    def init(self):
        a = self.store.arrays
        self.force.force(a['pmid'], a['disp'], float(self.cosmo.Omega_m), a['acc'])

    def step(self):
        from .nbody import _kick_drift
        i = self.i
        k1, d, k2 = self.inner.factors(i)
        if not self.pre:
            with TIMERS('kick_drift'):
                _kick_drift(self.store.ptcl, k1, d, True, True)
        a = self.store.arrays
        nxt = None
        if i + 1 < self.nsteps:
            k1n, dn, _ = self.inner.factors(i + 1)
            nxt = (k1n, dn)
        self.force.force(a['pmid'], a['disp'], float(self.cosmo.Omega_m), a['acc'], a['vel'], k2, next_kd=nxt,
                         sweep=self.store.sweep)
        self.pre = nxt is not None
        self.i += 1
        every = self.conf.reorder_every
        # (the gather pass has already applied drift self.i; the next one is what the forces to come will see)
        pred = 0.5 * (every - 1) * self.inner.factors(self.i + 1)[1] if (self.pre and self.i + 1 < self.nsteps) else 0.0
        self.store.maybe_reorder(sync_max=self.comm.allreduce_max, predict=pred)


def nbody_adj_slab(ptcl, ptcl_cot, cosmo, conf, comm, reverse=False, force=None, _a_nbody=None):
    """``nbody_adj`` (pmwd/nbody.py:251-260) on this rank's slab: returns
    ``(ptcl, ptcl_cot, cosmo_cot)`` with ``cosmo_cot`` already summed over ranks."""
    from .nbody import nbody_adj
    force = force or SlabForce(conf, comm)
    return nbody_adj(ptcl, ptcl_cot, None, cosmo, conf, reverse=reverse, _slab=force, _a_nbody=_a_nbody)


def nbody_step_slab(a_prev, a_next, ptcl, cosmo, conf, comm, force=None):
    """``nbody_step`` (pmwd/nbody.py:204-212) on this rank's slab: functional, returns new
    Particles (inputs untouched)."""
    from .nbody import _Store, _owned
    force = force or SlabForce(conf, comm)
    with torch.no_grad():
        p = _owned(ptcl, conf)
        store = _Store(conf, dict(pmid=p.pmid, disp=p.disp, vel=p.vel, acc=p.acc))
        step_slab(float(a_prev), float(a_next), store, cosmo, conf, force)
        a = store.arrays
    return Particles(conf, ptcl.pmid, a['disp'], vel=a['vel'], acc=a['acc'])


def nbody_step_slab_host(a_prev, a_next, host, cosmo, conf, comm, force=None, out=None, acc_resident=True):
    """``nbody_step`` on this rank's slab with the state in (pinned) HOST memory: the copy schedule of
    ``pmwd_b200.nbody_step_host`` (pmid resident, no re-upload of the accelerations the previous call wrote,
    displacement download under the force, acceleration download under the next call's uploads) around the
    collective slab force.  Collective: every rank calls it with its own Lagrangian slab."""
    from .nbody import nbody_step_host
    force = force or SlabForce(conf, comm)
    return nbody_step_host(a_prev, a_next, host, cosmo, conf, out=out, acc_resident=acc_resident, _slab=force)


def step_slab(a_prev, a_next, store, cosmo, conf, force):
    """One KDK step (default ``symp_splits``; nbody.py:121-140) on the store's arrays."""
    from .nbody import drift_factor, kick_factor, _f32, _kick_drift
    if tuple(conf.symp_splits) != ((0, 0.5), (1, 0.5)):
        raise NotImplementedError('the slab integrator implements the default KDK splitting')
    Om = float(cosmo.Omega_m)
    a_mid = a_prev * 0.5 + a_next * 0.5
    k1 = _f32(kick_factor(a_prev, a_prev, a_mid, cosmo, conf))
    d = _f32(drift_factor(a_mid, a_prev, a_next, cosmo, conf))
    k2 = _f32(kick_factor(a_next, a_mid, a_next, cosmo, conf))
    with TIMERS('kick_drift'):
        _kick_drift(store.ptcl, k1, d, True, True)
    a = store.arrays
    force.force(a['pmid'], a['disp'], Om, a['acc'], a['vel'], k2)


# ------------------------------------------------------------------------------------------
def init_process_group():
    if dist.is_initialized():
        return
    backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group(backend)


def run_bench(args):
    """bench.py at N > 1: weak scaling, per-GPU work fixed at n^3 particles / (2n)^3 cells."""
    import json
    import bench as B
    init_process_group()
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local)
    import pmwd_b200 as pm
    shape = B.rank_grid(args.n, world)
    conf = pm.Configuration(1., shape, mesh_shape=2, device=dev, reorder_every=args.reorder_every,
                            reorder_min_disp=args.reorder_min_disp)
    comm = SlabComm(conf)
    cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
    with torch.no_grad():
        white = white_noise_slab(0, conf, comm, dev, exact=False)
        ic = lpt_slab(white, cosmo, conf, comm)
        del white
    torch.cuda.empty_cache()
    from .nbody import _Store, _owned
    force = SlabForce(conf, comm)
    a = conf.a_nbody.tolist()
    nsched = len(a) - 1
    Om = float(cosmo.Omega_m)

    def fresh():
        p = _owned(ic, conf)
        st = _Store(conf, dict(pmid=p.pmid.clone(), disp=p.disp, vel=p.vel, acc=p.acc))
        sp = SlabStepper(st, a, cosmo, conf, comm, force)
        sp.init()
        return sp

    W, K = args.warmup, args.steps
    with torch.no_grad():
        stepper = fresh()
        for _ in range(W):
            if stepper.i == nsched:
                stepper = fresh()
            stepper.step()
        torch.cuda.synchronize(); dist.barrier()
        sampler = B.ClockSampler(local); sampler.start()
        TIMERS.on = True
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K):
            if stepper.i == nsched:
                stepper = fresh()
            stepper.step()
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms)
        launches = _lib.launch_count() - l0
        clocks = sampler.stop()
        phases = {k: round(v[0] / K, 3) for k, v in TIMERS.read().items()}
        TIMERS.on = False
    store = stepper.store
    assert torch.isfinite(store.arrays['disp']).all()
    Np = conf.ptcl_num

    # ---- reverse-time adjoint back over the section just timed (BASELINE config 5: forward + adjoint
    # on the slab decomposition; float64 dot products all-reduced once at the end)
    fwd_adjoint = adjoint = None
    if not getattr(args, 'no_adjoint', False):
        i_end = stepper.i
        ka = max(1, min(K, i_end))
        section = a[i_end - ka:i_end + 1]
        with torch.no_grad():
            disp, vel = store.lagrangian('disp', 'vel')
            final = Particles(conf, ic.pmid, disp, vel=vel)
            g = torch.Generator(device=dev).manual_seed(1 + rank)
            cot = Particles(conf, ic.pmid, torch.randn(disp.shape, device=dev, generator=g),
                            vel=torch.randn(disp.shape, device=dev, generator=g))
            del store, stepper, disp, vel
            torch.cuda.empty_cache()
            # warm-up (untimed): the adjoint over the last 3 steps -- first launches and the allocator's first
            # device allocations of the adjoint's buffers stay out of the timed region (cf. bench.py)
            nbody_adj_slab(final, cot, cosmo, conf, comm, force=force, _a_nbody=section[-min(4, len(section)):])
            TIMERS.read()
            TIMERS.on = True
            la0 = _lib.launch_count()
            torch.cuda.synchronize(); dist.barrier()
            e0.record()
            _, pc, cc = nbody_adj_slab(final, cot, cosmo, conf, comm, force=force, _a_nbody=section)
            e1.record()
            torch.cuda.synchronize(); dist.barrier()
            ams = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(ams, op=dist.ReduceOp.MAX)
            ams = float(ams)
            aphases = {k: round(v[0] / ka, 3) for k, v in TIMERS.read().items()}
            TIMERS.on = False
            launches_adj = _lib.launch_count() - la0
        assert torch.isfinite(pc.disp).all() and torch.isfinite(pc.vel).all()
        aper = ams / ka
        Nm_ = conf.mesh_size
        peak_, _src = B._peaks()
        adjoint = {'ms_per_step': aper, 'steps': ka, 'includes': 'nbody_adj init (one force_adj) + steps',
                   'particle_steps_per_sec': Np / (aper * 1e-3), 'adjoint_over_forward': aper / (ms / K),
                   'gpu_launches': launches_adj, 'phase_ms_per_step_rank0': aphases,
                   'step_frac': (312 * Np + 156 * Nm_) / world / aper / 1e6 / peak_,
                   'cosmo_cot_Omega_m': float(cc['Omega_m'])}
        fwd_adjoint = {'value': Np / ((aper + ms / K) * 1e-3),
                       'unit': 'particle-steps/s (one forward + one adjoint step)',
                       'ms_per_step': aper + ms / K, 'adjoint_ms_per_step': aper, 'forward_ms_per_step': ms / K,
                       'step_frac': (444 * Np + 224 * Nm_) / world / (aper + ms / K) / 1e6 / peak_}
        del final, cot, pc
        torch.cuda.empty_cache()
    store = stepper = None

    # ---- e2e: per-rank pinned host slabs -> device -> nbody_step_slab -> host, every step
    ke = max(1, min(K, args.e2e_steps))
    with torch.no_grad():
        p0 = fresh().store.ptcl
        host = {k: torch.empty(getattr(p0, k).shape, dtype=getattr(p0, k).dtype).pin_memory()
                for k in ('pmid', 'disp', 'vel', 'acc')}
        for k in host:
            host[k].copy_(getattr(p0, k))
        del p0
        torch.cuda.empty_cache()
        # pmid stays resident after its first upload, acc is not re-uploaded (device mirror of the array the
        # previous step wrote): disp + vel up, disp + vel + acc down, every step
        h2d = sum(host[k].numel() * host[k].element_size() for k in ('disp', 'vel')) * world
        d2h = sum(host[k].numel() * host[k].element_size() for k in ('disp', 'vel', 'acc')) * world

        def e2e_step(j):
            nbody_step_slab_host(a[j], a[j + 1], host, cosmo, conf, comm, force, out=host)
        e2e_step(0)
        e2e_step(1 % nsched)                       # second warm-up: from here on acc is the array the API wrote
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for j in range(2, 2 + ke):
            e2e_step(j % nsched)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        ems = float(ems)
        from .nbody import nbody_host_release
        nbody_host_release()
    e2e = {'value': Np * ke / (ems * 1e-3), 'unit': 'particle-updates/s', 'steps': ke, 'ms_per_step': ems / ke,
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'api': 'pmwd_b200.dist.nbody_step_slab_host: per-rank pinned host slabs in and out every step (disp, vel up; '
                  'disp, vel, acc down; acc not re-uploaded; disp download under the force, acc download under the '
                  'next step\'s uploads)'}
    if rank == 0:
        Nm = conf.mesh_size
        peak, peak_src = B._peaks()
        line = {
            'metric': 'particle_updates_per_sec', 'value': Np * K / (ms * 1e-3),
            'unit': 'particle-updates/s', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': B.workload_config(args, world),
            'steps_per_sec': K / (ms * 1e-3), 'clocks': clocks, 'gpu_launches': launches,
            'halo_planes': force.h_alloc, 'phase_ms_per_step_rank0': phases,
            'roofline': {'bound': 'hbm', 'unit': 'GB/s', 'peak': peak, 'peak_source': peak_src,
                         'kernel': 'whole step (per GPU)', 'traffic': None,
                         'achieved': (132 * Np + 68 * Nm) / world / (ms / K) / 1e6,
                         'frac': (132 * Np + 68 * Nm) / world / (ms / K) / 1e6 / peak},
            'e2e': e2e, 'fwd_adjoint': fwd_adjoint, 'adjoint': adjoint, 'cpu_baseline': None,
        }
        getattr(args, '_emit', lambda l: print(json.dumps(l), flush=True))(line)
    dist.barrier()
