"""pmwd_b200: B200-native particle-mesh hot path behind pmwd's Python API
(mirror of ``pmwd/__init__.py:4-15``)."""
from .configuration import Configuration
from .cosmology import Cosmology, SimpleLCDM, Planck18, E2, H_deriv, Omega_m_a
from .boltzmann import (transfer_integ, transfer_fit, transfer, growth_integ, growth,
                        boltzmann, linear_power)
from .particles import Particles, ptcl_rpos
from .scatter import scatter
from .gather import gather
from .gravity import laplace, neg_grad, gravity
from .modes import white_noise, linear_modes
from .lpt import lpt
from .nbody import nbody, nbody_init, nbody_step, nbody_step_host, nbody_host_release, nbody_adj
from .pm_util import fftfreq, fftfwd, fftinv
from .spec_util import powspec
from . import _lib

__version__ = '0.1.0'
