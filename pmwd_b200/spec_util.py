"""Power spectrum in spherical bins with the reference's API (``pmwd/spec_util.py:10-147``).

The FFT is cuFFT (``pm_util.fftfwd``); |f_k|^2 (or f_k conj(g_k)), the sinc deconvolution, the
Hermitian multiplicities and the ``digitize`` binning are one CUDA pass (``pmwd_powspec_bin``)
accumulating in float64; the handful of per-bin divisions happen in float64 torch on the device.
The auto spectrum is differentiable w.r.t. the field (it is the typical objective): the VJP is one
weighting pass (``pmwd_powspec_weight``) and one unnormalised C2R transform.  A cross spectrum of
fields that require grad is rejected rather than silently detached.
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


def _bin_sums(fields, others, grid_shape, edges_t, bnum, right, deconv):
    """float64 sums per digitize bin: rows = (k N, Re P N, Im P N, N), columns 0..bnum."""
    dev = fields.device
    lib = _lib.lib()
    sums = torch.zeros((4, bnum + 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        for n in range(fields.shape[0]):                     # leading axes are summed (:112-113)
            fk = fftfwd(fields[n]).contiguous()
            gk = None if others is None else fftfwd(others[n]).contiguous()
            _lib.check(lib.pmwd_powspec_bin(
                _lib.stream_ptr(dev), _lib.shape_arr(grid_shape), _lib.ptr(fk), _lib.ptr(gk),
                int(deconv is not None), float(deconv or 0.), _lib.ptr(edges_t), bnum, int(right),
                _lib.ptr(sums)), 'pmwd_powspec_bin')
    return sums


class _AutoSpec(torch.autograd.Function):
    """``fields (n, X, Y, Z) -> (sum Re P N per bin)``; everything else of powspec is
    field independent.  VJP: see ``pmwd_powspec_weight``."""

    @staticmethod
    def forward(ctx, fields, grid_shape, edges_t, bnum, right, deconv):
        sums = _bin_sums(fields, None, grid_shape, edges_t, bnum, right, deconv)
        ctx.save_for_backward(fields, edges_t)
        ctx.meta = (grid_shape, bnum, right, deconv)
        ctx.mark_non_differentiable(sums)
        return sums[1].clone(), sums

    @staticmethod
    def backward(ctx, pbar, _):
        fields, edges_t = ctx.saved_tensors
        grid_shape, bnum, right, deconv = ctx.meta
        dev = fields.device
        lib = _lib.lib()
        wbin = pbar.to(torch.float64).contiguous()           # cotangent of sum_{k in b} N_k D_k |f_k|^2
        grad = torch.empty_like(fields)
        with torch.cuda.device(dev):
            for n in range(fields.shape[0]):
                fk = fftfwd(fields[n]).contiguous()
                _lib.check(lib.pmwd_powspec_weight(
                    _lib.stream_ptr(dev), _lib.shape_arr(grid_shape), _lib.ptr(fk),
                    int(deconv is not None), float(deconv or 0.), _lib.ptr(edges_t), bnum, int(right),
                    _lib.ptr(wbin), _lib.ptr(fk)), 'pmwd_powspec_weight')
                # d|f_k|^2/df(x) summed over the full spectrum = 2 x unnormalised C2R of w_k f_k
                grad[n] = 2 * torch.fft.irfftn(fk, s=grid_shape, norm='forward')
        return grad, None, None, None, None, None


def powspec(f, spacing, bins=1j / 3, g=None, deconv=None, cut_zero=True, cut_nyq=True):
    """Auto or cross power spectrum in 3-D averaged in spherical bins
    (``pmwd/spec_util.py:50-147``).  Returns ``(k, P, N, bins)`` as float64 tensors
    (``P`` complex128 for a cross spectrum) on the device of ``f``."""
    f = torch.as_tensor(f)
    _lib.require_cuda(f)
    if g is not None:
        g = torch.as_tensor(g, device=f.device)
        if f.shape != g.shape:
            raise ValueError(f'shape mismatch: {tuple(f.shape)} != {tuple(g.shape)}')
        if f.requires_grad or g.requires_grad:
            raise NotImplementedError('the cross spectrum is forward only; detach the fields first')
    if f.ndim < 3:
        raise ValueError('the field needs at least 3 axes')
    grid_shape = tuple(f.shape[-3:])
    bnum, bcut, edges, right = _getbins(grid_shape, bins, cut_nyq)

    dev = f.device
    edges_t = torch.tensor(edges, dtype=torch.float64, device=dev)
    fields = f.reshape((-1,) + grid_shape).to(torch.float32)
    if g is None:
        pre, sums = _AutoSpec.apply(fields, grid_shape, edges_t, bnum, right, deconv)
        P = pre[:bnum]
    else:
        others = g.reshape((-1,) + grid_shape).to(torch.float32)
        sums = _bin_sums(fields, others, grid_shape, edges_t, bnum, right, deconv)
        P = torch.complex(sums[1, :bnum], sums[2, :bnum])
    # k N and N are field independent: the kernel accumulated them once per leading-axis field,
    # the reference counts every mode once (spec_util.py:112-113 sums P only)
    nfield = fields.shape[0]
    ksum, num = sums[0, :bnum] / nfield, sums[3, :bnum] / nfield

    lo = int(bool(cut_zero))
    k, P, N = ksum[lo:bcut], P[lo:bcut], num[lo:bcut]
    k = k / N * (2 * math.pi / spacing)
    P = P / N * (spacing ** 3 / math.prod(grid_shape))
    return k, P, N, edges_t[:bcut] * (2 * math.pi / spacing)
