"""Oracle configuration: restates the hot-path fields and derived properties of
``pmwd/configuration.py:101-320`` (the reference's ``Configuration``) without
JAX / mcfit.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import math

import numpy as np


class Conf:
    """Fields and defaults: ``pmwd/configuration.py:101-146``."""

    def __init__(self, ptcl_spacing, ptcl_grid_shape, mesh_shape=1,
                 float_dtype=np.float32, pmid_dtype=np.int16,
                 lpt_order=2, a_start=1 / 64, a_stop=1., a_lpt_maxstep=1 / 128,
                 a_nbody_maxstep=1 / 64, symp_splits=((0, 0.5), (1, 0.5)),
                 chunk_size=2 ** 24, k_pivot_Mpc=0.05, T_cmb=2.7255,
                 transfer_fit_nowiggle=False):
        self.ptcl_spacing = float(ptcl_spacing)
        self.ptcl_grid_shape = tuple(int(s) for s in ptcl_grid_shape)
        # configuration.py:152-161
        if isinstance(mesh_shape, (int, float)):
            mesh_shape = tuple(round(s * mesh_shape) for s in self.ptcl_grid_shape)
        self.mesh_shape = tuple(int(s) for s in mesh_shape)
        if len(self.ptcl_grid_shape) != len(self.mesh_shape):
            raise ValueError('particle and mesh grid dimensions differ')
        if any(sm < sp for sp, sm in zip(self.ptcl_grid_shape, self.mesh_shape)):
            raise ValueError('mesh grid cannot be smaller than particle grid')
        if any(self.ptcl_grid_shape[0] * sm != self.mesh_shape[0] * sp
               for sp, sm in zip(self.ptcl_grid_shape[1:], self.mesh_shape[1:])):
            raise ValueError('particle and mesh grid aspect ratios differ')
        self.float_dtype = np.dtype(float_dtype)
        self.pmid_dtype = np.dtype(pmid_dtype)
        self.cosmo_dtype = np.dtype(np.float64)
        self.lpt_order = lpt_order
        self.a_start = a_start
        self.a_stop = a_stop
        self.a_lpt_maxstep = a_lpt_maxstep
        self.a_nbody_maxstep = a_nbody_maxstep
        self.symp_splits = symp_splits
        self.chunk_size = chunk_size
        self.k_pivot_Mpc = k_pivot_Mpc
        self.T_cmb = T_cmb
        self.transfer_fit_nowiggle = transfer_fit_nowiggle
        # configuration.py:181-187
        if any(len(s) != 2 for s in symp_splits):
            raise ValueError(f'symp_splits={symp_splits} not supported')
        ssum = tuple(sum(s) for s in zip(*symp_splits))
        if ssum != (1, 1):
            raise ValueError(f'sum of symplectic splits = {ssum} != (1, 1)')
        # growth tolerances: configuration.py:180-185
        self.growth_rtol = self.growth_atol = math.sqrt(np.finfo(np.float64).eps)

    # constants: configuration.py:113-124
    M_sun_SI = 1.98847e30
    Mpc_SI = 3.0856775815e22
    H_0_SI = 1e5 / Mpc_SI
    c_SI = 299792458
    G_SI = 6.67430e-11
    M = 1e10 * M_sun_SI
    L = Mpc_SI
    T = 1 / H_0_SI
    transfer_lgk_min = -4
    transfer_lgk_max = 3
    transfer_lgk_maxstep = 1 / 128

    @property
    def dim(self):  # configuration.py:198-201
        return len(self.ptcl_grid_shape)

    @property
    def ptcl_cell_vol(self):  # :203-206
        return self.ptcl_spacing ** self.dim

    @property
    def ptcl_num(self):  # :208-212
        return int(np.prod(self.ptcl_grid_shape, dtype=np.int64))

    @property
    def box_size(self):  # :214-217
        return tuple(self.ptcl_spacing * s for s in self.ptcl_grid_shape)

    @property
    def box_vol(self):  # :219-223
        return float(np.prod(self.box_size))

    @property
    def cell_size(self):  # :225-228
        return self.ptcl_spacing * self.ptcl_grid_shape[0] / self.mesh_shape[0]

    @property
    def mesh_size(self):  # :235-239
        return int(np.prod(self.mesh_shape, dtype=np.int64))

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
    def transfer_k_num(self):  # :266-270
        return 1 + math.ceil((self.transfer_lgk_max - self.transfer_lgk_min)
                             / self.transfer_lgk_maxstep) + 1

    @property
    def transfer_k(self):  # :278-283
        k = np.logspace(self.transfer_lgk_min, self.transfer_lgk_max,
                        num=self.transfer_k_num - 1, dtype=np.float64)
        return np.concatenate((np.array([0.]), k))

    @property
    def a_lpt_num(self):  # :285-288
        return math.ceil(self.a_start / self.a_lpt_maxstep)

    @property
    def a_lpt_step(self):
        return self.a_start / self.a_lpt_num

    @property
    def a_nbody_num(self):  # :295-298
        return math.ceil((self.a_stop - self.a_start) / self.a_nbody_maxstep)

    @property
    def a_lpt(self):  # :305-309
        return np.linspace(0, self.a_start, num=self.a_lpt_num + 1, dtype=np.float64)

    @property
    def a_nbody(self):  # :311-315
        return np.linspace(self.a_start, self.a_stop, num=1 + self.a_nbody_num,
                           dtype=np.float64)

    @property
    def growth_a(self):  # :317-320
        return np.concatenate((self.a_lpt, self.a_nbody[1:]))
