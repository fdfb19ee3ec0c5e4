"""Oracle power-spectrum estimator: restates ``pmwd/spec_util.py:10-147``.
It is the parity metric ("P(k) within 0.1 % at all k").  TEST INFRASTRUCTURE ONLY.
"""
import math

import numpy as np

from .gravity import fftfreq, fftfwd


def _getbins(grid_shape, bins, cut_nyq):
    """``pmwd/spec_util.py:10-47``."""
    kfun = 1 / max(grid_shape)
    knyq = 0.5
    kmax = knyq * math.sqrt(3)
    if isinstance(bins, (int, float)):
        bins *= kfun
        bnum = 1 + math.ceil(kmax / bins)
        bcut = 1 + math.ceil(knyq / bins) if cut_nyq else bnum
        bins = bins * np.arange(bnum)
        right = True
    elif isinstance(bins, complex):
        kmaxable = all(s % 2 == 0 for s in grid_shape)
        bnum = 1 + math.ceil(math.log2(kmax / kfun) / bins.imag) + kmaxable
        bcut = 1 + math.ceil(math.log2(knyq / kfun) / bins.imag) if cut_nyq else bnum
        bins = kfun * 2 ** (bins.imag * np.arange(bnum))
        right = False
    elif isinstance(bins, tuple):
        if bins[0] != 0:
            raise ValueError(f'{bins=} must starts from 0')
        bnum = len(bins)
        if cut_nyq:
            for bcut, edge in enumerate(bins, start=1):
                if edge >= knyq:
                    break
        else:
            bcut = bnum
        bins = np.asarray(bins)
        right = True
    else:
        raise ValueError(f'{bins=} not supported')
    return bnum, bcut, bins, right


def powspec(f, spacing, bins=1j / 3, g=None, deconv=None, cut_zero=True, cut_nyq=True):
    """``pmwd/spec_util.py:50-147``: returns ``(k, P, N, bins)`` in float64."""
    f = np.asarray(f)
    if g is not None and f.shape != np.shape(g):
        raise ValueError(f'shape mismatch: {f.shape} != {np.shape(g)}')
    grid_shape = f.shape[-3:]
    bnum, bcut, bins, right = _getbins(grid_shape, bins, cut_nyq)

    last_three = tuple(range(-3, 0))
    f = fftfwd(f, axes=last_three)
    if g is None:
        P = f.real ** 2 + f.imag ** 2
    else:
        g = fftfwd(np.asarray(g), axes=last_three)
        P = f * g.conj()
    if P.ndim > 3:
        P = P.sum(tuple(range(P.ndim - 3)))

    kvec = fftfreq(grid_shape, None, dtype=P.real.dtype)
    k = np.sqrt(sum(k ** 2 for k in kvec))
    if deconv is not None:
        for kk in kvec:
            P = P * np.sinc(kk) ** -deconv

    N = np.full(P.shape, 2, dtype=np.uint32)
    N[..., 0] = 1
    if grid_shape[-1] % 2 == 0:
        N[..., -1] = 1

    k = k.ravel()
    P = P.ravel()
    N = N.ravel()
    b = np.digitize(k, bins, right=right)
    k = (k * N).astype(np.float64)
    P = (P * N).astype(np.float64 if not np.iscomplexobj(P) else np.complex128)
    k = np.bincount(b, weights=k, minlength=bnum)[:bnum]
    if np.iscomplexobj(P):
        P = (np.bincount(b, weights=P.real, minlength=bnum)[:bnum]
             + 1j * np.bincount(b, weights=P.imag, minlength=bnum)[:bnum])
    else:
        P = np.bincount(b, weights=P, minlength=bnum)[:bnum]
    N = np.bincount(b, weights=N, minlength=bnum)[:bnum]

    k = k[cut_zero:bcut]
    P = P[cut_zero:bcut]
    N = N[cut_zero:bcut]
    bins = bins[:bcut]

    with np.errstate(invalid='ignore', divide='ignore'):
        k = k / N
        P = P / N
    k = k * (2 * math.pi / spacing)
    bins = bins * (2 * math.pi / spacing)
    P = P * (spacing ** 3 / math.prod(grid_shape))
    return k, P, N, bins
