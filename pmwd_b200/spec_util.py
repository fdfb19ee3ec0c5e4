"""Power spectrum in spherical bins with the reference's API (``pmwd/spec_util.py:10-147``).

The FFT is cuFFT (``pm_util.fftfwd``); |f_k|^2 (or f_k conj(g_k)), the sinc deconvolution, the
Hermitian multiplicities and the ``digitize`` binning are one CUDA pass (``pmwd_powspec_bin``)
accumulating in float64; the handful of per-bin divisions happen in float64 torch on the device.
Forward only: the estimator is the parity metric of the N-body path; a field that requires grad
is rejected rather than silently detached.
"""
import ctypes as C
import math

import torch

from . import _lib
from .pm_util import fftfwd


def _getbins(grid_shape, bins, cut_nyq):
    """Bin edges in cycles per grid unit (``pmwd/spec_util.py:10-47``): a number = linear bins
    of that many fundamentals, an imaginary number = log2 spacing, a tuple = explicit edges.
    The first bin holds only the DC mode.  Returns ``(bnum, bcut, edges, right)``."""
    kfun = 1 / max(grid_shape)
    knyq = 0.5
    kmax = knyq * math.sqrt(3)
    if isinstance(bins, complex):
        step = bins.imag
        extra = all(s % 2 == 0 for s in grid_shape)
        bnum = 1 + math.ceil(math.log2(kmax / kfun) / step) + extra
        bcut = 1 + math.ceil(math.log2(knyq / kfun) / step) if cut_nyq else bnum
        edges = [kfun * 2 ** (step * b) for b in range(bnum)]
        return bnum, bcut, edges, False
    if isinstance(bins, (int, float)):
        width = bins * kfun
        bnum = 1 + math.ceil(kmax / width)
        bcut = 1 + math.ceil(knyq / width) if cut_nyq else bnum
        return bnum, bcut, [width * b for b in range(bnum)], True
    if isinstance(bins, tuple):
        if bins[0] != 0:
            raise ValueError(f'{bins=} must starts from 0')
        bnum = bcut = len(bins)
        if cut_nyq:
            for bcut, edge in enumerate(bins, start=1):
                if edge >= knyq:
                    break
        return bnum, bcut, [float(b) for b in bins], True
    raise ValueError(f'{bins=} not supported')


def powspec(f, spacing, bins=1j / 3, g=None, deconv=None, cut_zero=True, cut_nyq=True):
    """Auto or cross power spectrum in 3-D averaged in spherical bins
    (``pmwd/spec_util.py:50-147``).  Returns ``(k, P, N, bins)`` as float64 tensors
    (``P`` complex128 for a cross spectrum) on the device of ``f``."""
    f = torch.as_tensor(f)
    _lib.require_cuda(f)
    if f.requires_grad or (isinstance(g, torch.Tensor) and g.requires_grad):
        raise NotImplementedError('powspec is forward only; detach the field first')
    if g is not None:
        g = torch.as_tensor(g, device=f.device)
        if f.shape != g.shape:
            raise ValueError(f'shape mismatch: {tuple(f.shape)} != {tuple(g.shape)}')
    if f.ndim < 3:
        raise ValueError('the field needs at least 3 axes')
    grid_shape = tuple(f.shape[-3:])
    bnum, bcut, edges, right = _getbins(grid_shape, bins, cut_nyq)

    dev = f.device
    lib = _lib.lib()
    edges_t = torch.tensor(edges, dtype=torch.float64, device=dev)
    sums = torch.zeros((4, bnum + 1), dtype=torch.float64, device=dev)
    fields = f.reshape((-1,) + grid_shape).to(torch.float32)
    others = None if g is None else g.reshape((-1,) + grid_shape).to(torch.float32)
    with torch.cuda.device(dev):
        for n in range(fields.shape[0]):                     # leading axes are summed (:112-113)
            fk = fftfwd(fields[n]).contiguous()
            gk = None if others is None else fftfwd(others[n]).contiguous()
            _lib.check(lib.pmwd_powspec_bin(
                _lib.stream_ptr(dev), _lib.shape_arr(grid_shape), _lib.ptr(fk), _lib.ptr(gk),
                int(deconv is not None), float(deconv or 0.), _lib.ptr(edges_t), bnum, int(right),
                _lib.ptr(sums)), 'pmwd_powspec_bin')
    ksum, pre, pim, num = (sums[i, :bnum] for i in range(4))
    P = pre if g is None else torch.complex(pre, pim)

    lo = int(bool(cut_zero))
    k, P, N = ksum[lo:bcut], P[lo:bcut], num[lo:bcut]
    k = k / N * (2 * math.pi / spacing)
    P = P / N * (spacing ** 3 / math.prod(grid_shape))
    return k, P, N, edges_t[:bcut] * (2 * math.pi / spacing)
