"""Cosmology parameters and background functions on the host, float64 torch tensors
with autograd (stand-in for the reference's JAX pytree, ``pmwd/cosmology.py:15-269``;
JAX is not available in this image -- SURVEY.md H1).  Off the hot path: these are
scalars that feed the kick/drift factors.
"""
import dataclasses
from typing import ClassVar, Optional

import torch

from .configuration import Configuration

_PARAMS = ('A_s_1e9', 'n_s', 'Omega_m', 'Omega_b', 'h', 'Omega_k_', 'w_0_', 'w_a_')
_TABLES = ('transfer', 'growth', 'varlin')


def _as64(x, dtype):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.to(dtype=dtype, device='cpu')
    return torch.tensor(x, dtype=dtype)


@dataclasses.dataclass(frozen=True, eq=False)
class Cosmology:
    """Fields of ``pmwd/cosmology.py:53-72``.  Leaves are 0-d float64 CPU tensors (they may
    carry ``requires_grad``); ``+ - *`` act leafwise like the reference's cotangent
    algebra (``cosmology.py:83-93``)."""

    conf: Configuration = dataclasses.field(repr=False)

    A_s_1e9: torch.Tensor
    n_s: torch.Tensor
    Omega_m: torch.Tensor
    Omega_b: torch.Tensor
    h: torch.Tensor

    Omega_k_: Optional[torch.Tensor] = None
    Omega_k_fixed: ClassVar[float] = 0
    w_0_: Optional[torch.Tensor] = None
    w_0_fixed: ClassVar[float] = -1
    w_a_: Optional[torch.Tensor] = None
    w_a_fixed: ClassVar[float] = 0

    transfer: Optional[torch.Tensor] = None
    growth: Optional[torch.Tensor] = None
    varlin: Optional[torch.Tensor] = None

    def __post_init__(self):
        dtype = self.conf.cosmo_dtype
        for name in _PARAMS + _TABLES:
            object.__setattr__(self, name, _as64(getattr(self, name), dtype))

    def replace(self, **changes):
        return dataclasses.replace(self, **changes)

    def leaves(self):
        """Differentiable children, name -> tensor (None leaves skipped)."""
        return {n: getattr(self, n) for n in _PARAMS + _TABLES if getattr(self, n) is not None}

    def _map(self, fn, other=None):
        kw = {}
        for n in _PARAMS + _TABLES:
            x = getattr(self, n)
            if x is None:
                continue
            kw[n] = fn(x) if other is None else fn(x, getattr(other, n))
        return self.replace(**kw)

    def __add__(self, other):
        return self._map(torch.add, other)

    def __sub__(self, other):
        return self._map(torch.sub, other)

    def __mul__(self, other):
        return self._map(lambda x: x * other)

    __rmul__ = __mul__

    # cosmology.py:112-150
    @property
    def k_pivot(self):
        return self.conf.k_pivot_Mpc / (self.h * self.conf.Mpc_SI) * self.conf.L

    @property
    def A_s(self):
        return self.A_s_1e9 * 1e-9

    @property
    def Omega_c(self):
        return self.Omega_m - self.Omega_b

    @property
    def Omega_k(self):
        return self.Omega_k_fixed if self.Omega_k_ is None else self.Omega_k_

    @property
    def Omega_de(self):
        return 1 - (self.Omega_m + self.Omega_k)

    @property
    def w_0(self):
        return self.w_0_fixed if self.w_0_ is None else self.w_0_

    @property
    def w_a(self):
        return self.w_a_fixed if self.w_a_ is None else self.w_a_

    @property
    def ptcl_mass(self):
        return self.conf.rho_crit * self.Omega_m * self.conf.ptcl_cell_vol


def SimpleLCDM(conf, **kwargs):
    """``pmwd/cosmology.py:164-171``."""
    p = dict(A_s_1e9=2.0, n_s=0.96, Omega_m=0.3, Omega_b=0.05, h=0.7)
    p.update(kwargs)
    return Cosmology(conf, **p)


def Planck18(conf, **kwargs):
    """``pmwd/cosmology.py:174-181``."""
    p = dict(A_s_1e9=2.105, n_s=0.9665, Omega_m=0.3111, Omega_b=0.04897, h=0.6766)
    p.update(kwargs)
    return Cosmology(conf, **p)


def _t(a, cosmo):
    return torch.as_tensor(a, dtype=cosmo.conf.cosmo_dtype)


def E2(a, cosmo):
    """``pmwd/cosmology.py:185-219``."""
    a = _t(a, cosmo)
    de_a = a ** (-3 * (1 + cosmo.w_0 + cosmo.w_a)) * torch.exp(-3 * cosmo.w_a * (1 - a))
    return cosmo.Omega_m * a ** -3 + cosmo.Omega_k * a ** -2 + cosmo.Omega_de * de_a


def H_deriv(a, cosmo):
    """dlnH/dlna, ``pmwd/cosmology.py:222-242`` (the reference differentiates ``E2`` with
    JAX AD; the derivative is written out here and stays differentiable w.r.t. cosmo)."""
    a = _t(a, cosmo)
    p = -3 * (1 + cosmo.w_0 + cosmo.w_a)
    de_a = a ** p * torch.exp(-3 * cosmo.w_a * (1 - a))
    dde = de_a * (p / a + 3 * cosmo.w_a)
    dE2 = -3 * cosmo.Omega_m * a ** -4 - 2 * cosmo.Omega_k * a ** -3 + cosmo.Omega_de * dde
    return 0.5 * a * dE2 / E2(a, cosmo)


def Omega_m_a(a, cosmo):
    """``pmwd/cosmology.py:245-269``."""
    a = _t(a, cosmo)
    return cosmo.Omega_m / (a ** 3 * E2(a, cosmo))
