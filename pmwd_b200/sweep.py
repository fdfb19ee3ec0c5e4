"""Host side of the tiled ("pencil sweep") CIC deposit, ``csrc/scatter_sweep.cu``.

The reference's ``_scatter`` (``pmwd/scatter.py:33-83``) is one XLA scatter-add over all particles.
Inside the integrator the deposit instead goes through shared-memory mesh tiles: the particle
storage is kept sorted by (y-tile, z-tile, x-plane, y, z) and a table gives the particle range of
every (tile, plane).  This module owns that table and the scratch area (work counters + straggler
list) and hands them to the C library as a ``pmwd_sweep``.
"""
import ctypes as C
import os

import torch

from . import _lib

# planes per x segment of a work item: 64 (deposit 3.41-3.48 ms; 128: 3.52 ms, profiles/r02_sweep_tiles.txt)
LX = int(os.environ.get('PMWD_SWEEP_LX', '64'))


def enabled(conf):
    if not getattr(conf, 'scatter_tiled', True) or os.environ.get('PMWD_SWEEP', '1') == '0':
        return False
    if conf.scatter_mode == 'deterministic':
        return os.environ.get('PMWD_SWEEP_DET', '1') != '0'
    return conf.scatter_mode == 'atomic'


class SweepState:
    """Table + scratch of one particle store for one mesh descriptor."""

    def __init__(self, desc, dev, lx=LX, det=False):
        lib = _lib.lib()
        self.det = bool(det)       # also hold the halo arrays of the deterministic variant
        self.halo = None
        ty, bw = C.c_int32(0), C.c_int32(0)
        ok = lib.pmwd_sweep_pick(C.byref(desc), C.byref(ty), C.byref(bw))
        self.ty, self.bw = (int(ty.value), int(bw.value)) if ok else (0, 0)
        self.ok = False
        self.dev = dev
        if self.ty <= 0:
            return
        self.lx = lx
        self.table = self.scratch = None
        self.status = torch.zeros(2, dtype=torch.int64, device=dev)
        self._struct = None
        self.fit(desc)

    def fit(self, desc):
        """Size the table / scratch for this descriptor (slab runs: the planes held change with the halo
        width, the particle count with every migration); anything that has to be re-allocated invalidates
        the state until the next :meth:`build`."""
        lib = _lib.lib()
        self.nx_ext = int(desc.mesh_shape[0])
        self.n = int(desc.ptcl_num)
        tb = lib.pmwd_sweep_table_bytes(C.byref(desc), self.ty, self.bw) // 4
        if self.table is None or self.table.numel() != tb:
            self.table = torch.zeros(tb, dtype=torch.int32, device=self.dev)
            self.ok = False
        sb = lib.pmwd_sweep_scratch_bytes(C.byref(desc))
        if self.scratch is None or self.scratch.numel() < sb:
            self.scratch = torch.empty(sb + sb // 8, dtype=torch.uint8, device=self.dev)    # room to grow
            self.scratch[:64].zero_()      # counters (incl. the sticky one of the deterministic deposit)
            self.ok = False
        if self.det:
            hb = lib.pmwd_sweep_det_halo_bytes(C.byref(desc), self.ty, self.bw)
            if hb == 0:
                self.det = False
            elif self.halo is None or self.halo.numel() < hb:
                self.halo = torch.empty(hb, dtype=torch.uint8, device=self.dev)
                self.ok = False

    def _make_struct(self, desc):
        s = _lib.Sweep()
        s.table = self.table.data_ptr()
        s.ty, s.bw, s.lx = self.ty, self.bw, self.lx
        s.nx_ext = int(desc.mesh_shape[0])
        a1 = float(torch.tensor(desc.cell_size, dtype=torch.float32))
        s.xoff = int(desc.offset[0] // a1) % int(desc.wrap_shape[0])
        s.scratch = self.scratch.data_ptr()
        s.scratch_bytes = self.scratch.numel()
        if self.det and self.halo is not None:
            s.det_halo = self.halo.data_ptr()
            s.det_halo_bytes = self.halo.numel()
        self._struct = s

    def build(self, desc, keys_ptr, check=False):
        """(Re)build the table from the keys of the sort that has just ordered the storage.
        With ``check`` the device-side validation is read back (one synchronisation); an invalid
        table leaves the state unusable (``ok`` False) and the RED kernel in charge."""
        if self.ty <= 0:
            return False
        lib = _lib.lib()
        self.fit(desc)
        with torch.cuda.device(self.dev):
            _lib.check(lib.pmwd_sweep_table(_lib.stream_ptr(self.dev), C.byref(desc), self.ty, self.bw, keys_ptr,
                                            _lib.ptr(self.table), _lib.ptr(self.status)), 'pmwd_sweep_table')
        self.ok = True
        if check:
            st = self.status.cpu()
            bad = int(st[0]) & 0xffffffff
            total = int(st[1])
            self.ok = bad == 0 and total == self.n
        if self.ok:
            self._make_struct(desc)
        return self.ok

    def arg(self):
        """``const pmwd_sweep*`` for the C calls (NULL when unusable)."""
        if not self.ok or self._struct is None:
            return None
        return C.byref(self._struct)

    def usable(self, desc):
        """True if the C library would run the tiled kernels for this descriptor (slab runs: the table
        was built for the same planes)."""
        a = self.arg()
        return a is not None and bool(_lib.lib().pmwd_sweep_usable(C.byref(desc), a))

    def det_violations(self):
        """Particles a deterministic deposit had to hand to the order-dependent straggler path (0 for storage
        that was re-sorted before every deposit).  Synchronises."""
        if not self.ok or not self.det:
            return 0
        with torch.cuda.device(self.dev):
            return int(_lib.lib().pmwd_sweep_det_violations(_lib.stream_ptr(self.dev), C.byref(self._struct)))

    def stragglers(self):
        if not self.ok:
            return -1
        with torch.cuda.device(self.dev):
            return int(_lib.lib().pmwd_sweep_last_stragglers(_lib.stream_ptr(self.dev), C.byref(self._struct)))
