"""Configuration: field-compatible mirror of the reference's frozen dataclass
(``pmwd/configuration.py:21-320``) for the host side of the hot path.

Host arrays are torch tensors instead of JAX arrays (no JAX in this image); the
``sigma8`` / ``varlin`` machinery (``mcfit.TophatVar``, ``configuration.py:173-178``)
is off the hot path and not provided.
"""
import dataclasses
import math
from typing import Optional, Tuple, Union

import numpy as np
import torch


@dataclasses.dataclass(frozen=True)
class Configuration:
    """See ``pmwd/configuration.py:23-100`` for the meaning of each parameter."""

    ptcl_spacing: float
    ptcl_grid_shape: Tuple[int, ...]

    mesh_shape: Union[float, Tuple[int, ...]] = 1

    cosmo_dtype: torch.dtype = torch.float64
    pmid_dtype: torch.dtype = torch.int16
    float_dtype: torch.dtype = torch.float32

    k_pivot_Mpc: float = 0.05

    T_cmb: float = 2.7255

    # constants in SI units (configuration.py:113-118)
    M_sun_SI = 1.98847e30
    Mpc_SI = 3.0856775815e22
    H_0_SI = 1e5 / Mpc_SI
    c_SI = 299792458
    G_SI = 6.67430e-11

    # units (configuration.py:120-123)
    M: float = 1e10 * M_sun_SI
    L: float = Mpc_SI
    T: float = 1 / H_0_SI

    transfer_fit: bool = True
    transfer_fit_nowiggle: bool = False
    transfer_lgk_min: float = -4
    transfer_lgk_max: float = 3
    transfer_lgk_maxstep: float = 1 / 128

    growth_rtol: Optional[float] = None
    growth_atol: Optional[float] = None
    growth_inistep: Union[float, None, Tuple[Optional[float], Optional[float]]] = (1, None)

    lpt_order: int = 2

    a_start: float = 1 / 64
    a_stop: float = 1
    a_lpt_maxstep: float = 1 / 128
    a_nbody_maxstep: float = 1 / 64

    symp_splits: Tuple[Tuple[float, float], ...] = ((0, 0.5), (1, 0.5))

    chunk_size: int = 2 ** 24   # accepted for API compatibility; the kernels need no chunking

    # additions of this implementation (not in the reference)
    scatter_mode: str = 'atomic'          # 'atomic' | 'deterministic' (cell-sorted)
    reorder_every: int = 3                # re-sort the integrator's particle storage by mesh cell
                                          # every this many steps (0 = never; see csrc/reorder.cu)
    reorder_min_disp: float = 1.0         # ... once max |disp| exceeds this many cells
    scatter_tiled: bool = True            # integrator deposits through shared-memory mesh tiles (csrc/scatter_sweep.cu)
    device: Union[str, torch.device] = 'cuda'

    def __post_init__(self):
        set_ = object.__setattr__
        set_(self, 'ptcl_grid_shape', tuple(int(s) for s in self.ptcl_grid_shape))
        # configuration.py:152-161
        if isinstance(self.mesh_shape, (int, float)):
            set_(self, 'mesh_shape', tuple(round(s * self.mesh_shape)
                                           for s in self.ptcl_grid_shape))
        else:
            set_(self, 'mesh_shape', tuple(int(s) for s in self.mesh_shape))
        if len(self.ptcl_grid_shape) != len(self.mesh_shape):
            raise ValueError('particle and mesh grid dimensions differ')
        if any(sm < sp for sp, sm in zip(self.ptcl_grid_shape, self.mesh_shape)):
            raise ValueError('mesh grid cannot be smaller than particle grid')
        if any(self.ptcl_grid_shape[0] * sm != self.mesh_shape[0] * sp
               for sp, sm in zip(self.ptcl_grid_shape[1:], self.mesh_shape[1:])):
            raise ValueError('particle and mesh grid aspect ratios differ')

        # configuration.py:163-171
        if not self.cosmo_dtype.is_floating_point:
            raise ValueError('cosmo_dtype must be floating point numbers')
        if self.pmid_dtype not in (torch.int8, torch.int16, torch.int32):
            raise ValueError('pmid_dtype must be signed integers (int8, int16 or int32)')
        if self.float_dtype != torch.float32:
            raise ValueError('float_dtype: the B200 kernels are float32 '
                             '(the reference default); float64 is not provided')

        growth_tol = math.sqrt(torch.finfo(self.cosmo_dtype).eps)   # configuration.py:180-185
        if self.growth_rtol is None:
            set_(self, 'growth_rtol', growth_tol)
        if self.growth_atol is None:
            set_(self, 'growth_atol', growth_tol)

        # configuration.py:187-192
        if any(len(s) != 2 for s in self.symp_splits):
            raise ValueError(f'symp_splits={self.symp_splits} not supported')
        symp_splits_sum = tuple(sum(s) for s in zip(*self.symp_splits))
        if symp_splits_sum != (1, 1):
            raise ValueError(f'sum of symplectic splits = {symp_splits_sum} != (1, 1)')

        if self.scatter_mode not in ('atomic', 'deterministic'):
            raise ValueError(f'scatter_mode={self.scatter_mode} not supported')
        set_(self, 'device', torch.device(self.device))

    def replace(self, **changes):
        return dataclasses.replace(self, **changes)

    # derived properties: configuration.py:198-320
    @property
    def dim(self):
        return len(self.ptcl_grid_shape)

    @property
    def ptcl_cell_vol(self):
        return self.ptcl_spacing ** self.dim

    @property
    def ptcl_num(self):
        return math.prod(self.ptcl_grid_shape)

    @property
    def box_size(self):
        return tuple(self.ptcl_spacing * s for s in self.ptcl_grid_shape)

    @property
    def box_vol(self):
        return math.prod(self.box_size)

    @property
    def cell_size(self):
        return self.ptcl_spacing * self.ptcl_grid_shape[0] / self.mesh_shape[0]

    @property
    def cell_vol(self):
        return self.cell_size ** self.dim

    @property
    def mesh_size(self):
        return math.prod(self.mesh_shape)

    @property
    def V(self):
        return self.L / self.T

    @property
    def H_0(self):
        return self.H_0_SI * self.T

    @property
    def c(self):
        return self.c_SI / self.V

    @property
    def G(self):
        return self.G_SI * self.M / (self.L * self.V ** 2)

    @property
    def rho_crit(self):
        return 3 * self.H_0 ** 2 / (8 * math.pi * self.G)

    @property
    def transfer_k_num(self):
        return 1 + math.ceil((self.transfer_lgk_max - self.transfer_lgk_min)
                             / self.transfer_lgk_maxstep) + 1

    @property
    def transfer_lgk_step(self):
        return ((self.transfer_lgk_max - self.transfer_lgk_min)
                / (self.transfer_k_num - 2))

    @property
    def transfer_k(self):
        k = np.logspace(self.transfer_lgk_min, self.transfer_lgk_max,
                        num=self.transfer_k_num - 1, dtype=np.float64)
        return torch.from_numpy(np.concatenate((np.zeros(1), k))).to(self.cosmo_dtype)

    @property
    def a_lpt_num(self):
        return math.ceil(self.a_start / self.a_lpt_maxstep)

    @property
    def a_lpt_step(self):
        return self.a_start / self.a_lpt_num

    @property
    def a_nbody_num(self):
        return math.ceil((self.a_stop - self.a_start) / self.a_nbody_maxstep)

    @property
    def a_nbody_step(self):
        return (self.a_stop - self.a_start) / self.a_nbody_num

    @property
    def a_lpt(self):
        return torch.from_numpy(np.linspace(0, self.a_start, num=self.a_lpt_num + 1,
                                            dtype=np.float64)).to(self.cosmo_dtype)

    @property
    def a_nbody(self):
        return torch.from_numpy(np.linspace(self.a_start, self.a_stop,
                                            num=1 + self.a_nbody_num,
                                            dtype=np.float64)).to(self.cosmo_dtype)

    @property
    def growth_a(self):
        return torch.cat((self.a_lpt, self.a_nbody[1:]))
