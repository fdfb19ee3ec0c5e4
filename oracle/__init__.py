"""CPU oracle for the pmwd particle-mesh hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm
(eelregit/pmwd, ``/root/reference/pmwd/*.py``); every function cites the
reference file:line it follows.  It exists to *check* the CUDA product in
``pmwd_b200/``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``pmwd_b200/`` imports it, and the product never falls back to it.

Pinning status
--------------
The reference is pure JAX and cannot be imported in this image (no ``jax``, no
``mcfit``), so no reference-generated vectors exist.

* scatter / gather / enmesh: PINNED against the reference's own known-answer
  tests (``tests/pm_test.py:43-118``: centred-particle deposit ``n * 2**-dim``,
  mass conservation, gather of a uniform mesh) -- ``tests/test_oracle_pins.py``.
* VJPs: pinned structurally (analytic VJP == float64 finite differences and the
  dot-product identity), the oracle analogue of ``pmwd/test_util.py:88-149`` and
  ``tests/pm_test.py:121-171``.
* FFT / ``gravity`` output values, LPT and the growth ODE: the reference holds no
  golden values for them -> **parity unpinned** by the reference for those; we
  pin them ourselves with closed-form plane-wave tests, the EdS growth solution
  and the 3-digit soft known-answers of ``docs/examples/quickstart.ipynb``
  (sigma_disp after LPT), with tolerances stated in the tests.

Compiled twin
-------------
``oracle/cpm.c`` (+ ``oracle/cpm.py``) restates one forward KDK step in C with OpenMP, bit-identical
to the NumPy functions here when run on one thread (``tests/test_oracle_c.py``); it is what
``bench.py`` times as the CPU baseline.  Same rule: test and bench infrastructure only.
"""

from .conf import Conf
from .pm import (enmesh, scatter, gather, scatter_adj, gather_adj, gen_grid,
                 ptcl_pos)
from .gravity import (fftfreq, fftfwd, fftinv, laplace, neg_grad, gravity,
                      gravity_vjp, rho_to_force)
from .cosmo import (Cosmo, SimpleLCDM, E2, H_deriv, Omega_m_a, growth_integ,
                    growth, boltzmann, transfer_fit, linear_power)
from .nbody import (G_D, G_K, drift_factor, kick_factor, drift, kick,
                    integrate, nbody, nbody_init, nbody_step, nbody_adj,
                    factor_grads)
from .lpt import white_noise, linear_modes, lpt, lpt_vjp_modes
from .spec import powspec
