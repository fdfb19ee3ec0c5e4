"""Initial-condition modes (``pmwd/modes.py:12-117``).  One FFT, run once: off the hot
path; kept so that the reference's model pipeline (``boltzmann -> linear_modes -> lpt ->
nbody``) is a drop-in and so that gradients w.r.t. the white noise flow end to end.
"""
import numpy as np
import torch

from .boltzmann import linear_power
from .pm_util import fftfreq, fftfwd, fftinv


def white_noise(seed, conf, real=False, unit_abs=False, device=None):
    """``pmwd/modes.py:12-49``.  The stream is
    ``numpy.random.default_rng(seed).standard_normal(shape, float32)`` (synthetic inputs of
    SURVEY.md 8d); ``jax.random.normal(PRNGKey(seed))`` cannot be reproduced without JAX."""
    device = conf.device if device is None else device
    modes = np.random.default_rng(seed).standard_normal(conf.ptcl_grid_shape, dtype=np.float32)
    modes = torch.from_numpy(modes).to(conf.float_dtype).to(device)
    if real and not unit_abs:
        return modes
    modes = fftfwd(modes, norm='ortho')
    if unit_abs:
        modes = modes / modes.abs()
    if real:
        modes = fftinv(modes, shape=conf.ptcl_grid_shape, norm='ortho')
    return modes


class _SafeSqrt(torch.autograd.Function):
    """``_safe_sqrt`` (``pmwd/modes.py:52-64``)."""

    @staticmethod
    def forward(ctx, x):
        y = torch.sqrt(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, y_cot):
        (y,) = ctx.saved_tensors
        safe = torch.where(y != 0, y, torch.ones_like(y))
        return torch.where(y != 0, 0.5 / safe * y_cot, torch.zeros_like(y))


def linear_modes(modes, cosmo, conf, a=None, real=False):
    """``pmwd/modes.py:67-117``: ``modes * sqrt(P_lin(k) V)``."""
    modes = torch.as_tensor(modes)
    dev = modes.device
    kvec = fftfreq(conf.ptcl_grid_shape, conf.ptcl_spacing, dtype=conf.float_dtype, device=dev)
    k = torch.sqrt(sum(k ** 2 for k in kvec))
    if a is not None:
        a = float(np.float32(float(a)))
    Plin = linear_power(k, a, cosmo, conf)
    if not modes.is_complex():
        modes = fftfwd(modes, norm='ortho')
    modes = modes * _SafeSqrt.apply(Plin * conf.box_vol)
    if real:
        modes = fftinv(modes, shape=conf.ptcl_grid_shape, norm=conf.ptcl_spacing)
    return modes
