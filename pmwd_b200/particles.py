"""Particle state: mirror of ``pmwd/particles.py:18-209`` on torch tensors.

Layout contract for the kernels (unchanged from the reference): AoS rows, ``pmid``
int16 ``(N, dim)``, ``disp`` / ``vel`` / ``acc`` float32 ``(N, dim)``, particle order = C-order
ravel of the Lagrangian grid.
"""
import dataclasses
from typing import Any, Optional

import numpy as np
import torch

from .configuration import Configuration


@dataclasses.dataclass(frozen=True, eq=False)
class Particles:
    """``pmwd/particles.py:18-59``."""

    conf: Configuration = dataclasses.field(repr=False)

    pmid: torch.Tensor
    disp: torch.Tensor
    vel: Optional[torch.Tensor] = None
    acc: Optional[torch.Tensor] = None

    attr: Any = None

    def __post_init__(self):
        conf = self.conf
        for name in ('pmid', 'disp', 'vel', 'acc'):
            value = getattr(self, name)
            if value is None:
                continue
            dtype = conf.pmid_dtype if name == 'pmid' else conf.float_dtype
            value = torch.as_tensor(value)
            if value.dtype != dtype:
                value = value.to(dtype)
            object.__setattr__(self, name, value)

    def replace(self, **changes):
        return dataclasses.replace(self, **changes)

    def __len__(self):
        return len(self.pmid)

    def to(self, device):
        kw = {n: getattr(self, n).to(device) for n in ('pmid', 'disp', 'vel', 'acc')
              if getattr(self, n) is not None}
        return self.replace(**kw)

    @classmethod
    def from_pos(cls, conf, pos, wrap=True):
        """``pmwd/particles.py:81-107``."""
        pos = torch.as_tensor(pos)
        pmid = torch.round(pos / conf.cell_size)          # rint: round half to even
        disp = pos - pmid * conf.cell_size
        pmid = pmid.to(torch.int64)
        disp = disp.to(conf.float_dtype)
        if wrap:
            pmid = pmid % torch.tensor(conf.mesh_shape, dtype=torch.int64, device=pmid.device)
        return cls(conf, pmid.to(conf.pmid_dtype), disp)

    @classmethod
    def gen_grid(cls, conf, vel=False, acc=False, device=None):
        """``pmwd/particles.py:109-144``: uniform grid, Lagrangian (C-order) sequence."""
        device = conf.device if device is None else torch.device(device)
        pmid, disp = [], []
        for sp, sm in zip(conf.ptcl_grid_shape, conf.mesh_shape):
            p1 = np.rint(np.linspace(0, sm, num=sp, endpoint=False))
            d1 = np.arange(sp) * sm - p1.astype(int) * sp           # exact int arithmetic
            d1 = d1 * (conf.cell_size / sp)
            pmid.append(torch.from_numpy(p1).to(conf.pmid_dtype).to(device))
            disp.append(torch.from_numpy(d1).to(conf.float_dtype).to(device))
        pmid = torch.stack(torch.meshgrid(*pmid, indexing='ij'), dim=-1).reshape(-1, conf.dim)
        disp = torch.stack(torch.meshgrid(*disp, indexing='ij'), dim=-1).reshape(-1, conf.dim)
        v = torch.zeros_like(disp) if vel else None
        a = torch.zeros_like(disp) if acc else None
        return cls(conf, pmid.contiguous(), disp.contiguous(), vel=v, acc=a)

    def raveled_id(self, dtype=torch.int64, wrap=False):
        """``pmwd/particles.py:156-182``."""
        conf = self.conf
        pmid = self.pmid.to(torch.int64)
        if wrap:
            pmid = pmid % torch.tensor(conf.mesh_shape, dtype=torch.int64, device=pmid.device)
        strides = np.cumprod((1,) + conf.mesh_shape[:0:-1])[::-1]
        rid = sum(pmid[:, i] * int(s) for i, s in enumerate(strides))
        return rid.to(dtype)

    def pos(self, dtype=torch.float64, wrap=True):
        """``pmwd/particles.py:184-209``."""
        conf = self.conf
        pos = self.pmid.to(dtype)
        pos = pos * conf.cell_size
        pos = pos + self.disp.to(dtype)
        if wrap:
            pos = pos % torch.tensor(conf.box_size, dtype=dtype, device=pos.device)
        return pos


def ptcl_rpos(ptcl, ref, conf, wrap=True):
    """``pmwd/particles.py:259-288``."""
    if not isinstance(ref, Particles):
        ref = Particles.from_pos(conf, ref, wrap=False)
    rpos = (ptcl.pmid.to(torch.int32) - ref.pmid.to(torch.int32)).to(conf.float_dtype)
    rpos = rpos * conf.cell_size
    rpos = rpos + (ptcl.disp - ref.disp)
    if wrap:
        box = torch.tensor(conf.box_size, dtype=conf.float_dtype, device=rpos.device)
        rpos = rpos - torch.round(rpos / box) * box
    return rpos
