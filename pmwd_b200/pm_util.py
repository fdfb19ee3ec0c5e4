"""FFT helpers with the reference's conventions: ``pmwd/pm_util.py:159-344``
(``fftfreq``, ``fftfwd``, ``fftinv``).

These are the *general* wrappers (any axes / shape / norm) used by LPT and analysis code;
they call cuFFT through ``torch.fft`` and stay differentiable.  The force pipeline does not
go through them: it uses cached cuFFT plans inside ``libpmwd_b200.so`` (``pmwd_force``).
"""
import math

import numpy as np
import torch


class KVec(list):
    """List of sparse wavevector arrays that remembers the grid it was built for, so that
    ``laplace`` / ``neg_grad`` can run the fused CUDA kernels instead of broadcasting."""
    shape = None
    spacing = None


def fftfreq(shape, spacing, dtype=torch.float64, sparse=True, device=None):
    """``pmwd/pm_util.py:159-199``: angular wavevectors, computed in float64 then cast."""
    period = 1.0
    if spacing is not None:
        period = 2 * math.pi / spacing
    dim = len(shape)
    kvec = KVec()
    kvec.shape = tuple(int(s) for s in shape)
    kvec.spacing = spacing
    for axis, s in enumerate(shape):
        k = np.fft.rfftfreq(s) if axis == dim - 1 else np.fft.fftfreq(s)
        k = torch.from_numpy(k * period).to(dtype)
        if device is not None:
            k = k.to(device)
        if sparse:
            view = [1] * dim
            view[axis] = -1
            k = k.reshape(view)
        k._pmwd_meta = (kvec.shape, spacing, axis)
        kvec.append(k)
    if not sparse:
        dense = torch.meshgrid(*kvec, indexing='ij')
        out = KVec(dense)
        out.shape, out.spacing = kvec.shape, spacing
        return out
    return kvec


def _norm_dim(f, shape, axes):
    d = f.ndim
    if shape is not None:
        d = len(shape)
    if axes is not None:
        d = len(axes)
    return d


def fftfwd(f, shape=None, axes=None, norm=None):
    """``pmwd/pm_util.py:236-289``: ``rfftn``; a float ``norm`` is the grid spacing."""
    f = torch.as_tensor(f)
    if f.is_complex():
        raise ValueError('input field must be real')
    if norm in {None, 'backward', 'ortho', 'forward'}:
        return torch.fft.rfftn(f, s=shape, dim=axes, norm=norm or 'backward')
    d = _norm_dim(f, shape, axes)
    return norm ** d * torch.fft.rfftn(f, s=shape, dim=axes, norm='backward')


def fftinv(f, shape=None, axes=None, norm=None):
    """``pmwd/pm_util.py:292-344``: ``irfftn``; a float ``norm`` is the grid spacing."""
    f = torch.as_tensor(f)
    if not f.is_complex():
        raise ValueError('input field must be Hermitian complex')
    if axes is None and shape is not None:
        axes = tuple(range(-len(shape), 0))
    if norm in {None, 'backward', 'ortho', 'forward'}:
        return torch.fft.irfftn(f, s=shape, dim=axes, norm=norm or 'backward')
    d = _norm_dim(f, shape, axes)
    return norm ** -d * torch.fft.irfftn(f, s=shape, dim=axes, norm='backward')


def fft(f, shape=None, axes=None, norm=None):
    """``pmwd/pm_util.py:209-233``."""
    f = torch.as_tensor(f)
    if not f.is_complex():
        return fftfwd(f, shape, axes, norm)
    return fftinv(f, shape, axes, norm)
