"""Oracle CIC primitives: NumPy restatement of ``pmwd/pm_util.py:4-156``,
``pmwd/scatter.py:33-148``, ``pmwd/gather.py:33-142`` and
``pmwd/particles.py:109-144,184-209``.  TEST INFRASTRUCTURE ONLY.

Arithmetic follows the reference operation by operation in ``conf.float_dtype``
(float32 by default) so that it can serve as the parity target; pass float64
arrays and ``Conf(float_dtype=np.float64)`` for the high-precision twin.
"""
import numpy as np


def _chunks(ptcl_num, chunk_size):
    """Chunk boundaries, remainder first (``pmwd/pm_util.py:4-21``)."""
    chunk_size = ptcl_num if chunk_size is None else min(chunk_size, ptcl_num)
    rem = ptcl_num % chunk_size
    bounds = []
    if rem:
        bounds.append((0, rem))
    for c in range(ptcl_num // chunk_size):
        bounds.append((rem + c * chunk_size, rem + (c + 1) * chunk_size))
    return bounds


def enmesh(i1, d1, a1, s1, b12, a2, s2, grad):
    """Multilinear mesh indices and fractions (``pmwd/pm_util.py:33-156``).

    Returns ``i2 (num, 2**dim, dim)``, ``f2 (num, 2**dim)`` and, if ``grad``,
    ``f2_grad (num, 2**dim, dim)``.
    """
    i1 = np.asarray(i1)
    d1 = np.asarray(d1)
    fdt = d1.dtype
    idt = i1.dtype
    # pm_util.py:83-91
    a1 = np.float64(a1) if a2 is not None else fdt.type(a1)
    if s1 is not None:
        s1 = np.array(s1, dtype=idt)
    b12 = np.asarray(b12, dtype=np.float64)
    if a2 is not None:
        a2 = np.float64(a2)
    if s2 is not None:
        s2 = np.array(s2, dtype=idt)

    dim = i1.shape[1]
    # pm_util.py:94-97: neighbour n has offset bit j of n along axis j
    neighbors = (np.arange(2 ** dim, dtype=idt)[:, np.newaxis]
                 >> np.arange(dim, dtype=idt)) & 1

    if a2 is not None:
        # general float64 branch, pm_util.py:99-118
        P = i1 * a1 + d1 - b12
        P = P[:, np.newaxis]
        i2 = P + neighbors * a2
        if s1 is not None:
            L = s1 * a1
            i2 = i2 % L
        i2 = i2 // a2
        d2 = P - i2 * a2
        if s1 is not None:
            d2 = d2 - np.rint(d2 / L) * L
        i2 = i2.astype(idt)
        d2 = d2.astype(fdt)
        a2 = fdt.type(a2)
        d2 = d2 / a2
    else:
        # fast branch, pm_util.py:119-136
        i12, d12 = np.divmod(b12, np.float64(a1))
        i1 = i1 - i12.astype(idt)
        d1 = d1 - d12.astype(fdt)
        i1 = i1[:, np.newaxis]
        d1 = d1[:, np.newaxis]
        d1 = d1 / a1                       # float divide in float_dtype
        i2 = np.floor(d1).astype(idt)
        i2 = i2 + neighbors
        d2 = d1 - i2.astype(fdt)           # JAX promotes int16,float32 -> float32
        i2 = i2 + i1
        if s1 is not None:
            i2 = i2 % s1

    f2 = (1 - np.abs(d2)).astype(fdt)      # pm_util.py:138

    if s1 is None and s2 is not None:      # pm_util.py:140-141
        i2 = np.where(i2 < 0, s2, i2)

    if grad:
        # pm_util.py:143-152
        sign = np.sign(-d2)
        f2g = []
        for i in range(dim):
            not_i = tuple(range(i + 1, dim)) + tuple(range(0, i))
            f2g.append(sign[..., i] * _prod_last(f2[..., not_i]))
        f2g = np.stack(f2g, axis=-1).astype(fdt)
        f2 = _prod_last(f2)
        return i2, f2, f2g
    f2 = _prod_last(f2)
    return i2, f2


def _prod_last(x):
    """Sequential left-to-right product over the last axis (empty -> 1)."""
    out = np.ones(x.shape[:-1], dtype=x.dtype)
    for j in range(x.shape[-1]):
        out = out * x[..., j]
    return out


def _valid_linear(ind, spatial_shape):
    """Linear index + validity mask.  JAX ``.at[].add`` drops out-of-bounds
    updates and ``.at[].get(mode='drop', fill_value=0)`` fills them with 0; negative
    indices would wrap NumPy-style, which ``enmesh(s2=...)`` prevents by mapping
    them to ``s2`` (``pmwd/pm_util.py:140-141``)."""
    shape = np.asarray(spatial_shape, dtype=np.int64)
    ind = ind.astype(np.int64)
    neg = ind < 0
    ind = np.where(neg, ind + shape, ind)   # numpy/JAX negative index semantics
    valid = np.all((ind >= 0) & (ind < shape), axis=-1)
    ind = np.where(valid[..., np.newaxis], ind, 0)
    lin = np.ravel_multi_index(tuple(np.moveaxis(ind, -1, 0)), tuple(shape))
    return lin, valid


def _prep(pmid, conf, mesh_or_none, val, default_val):
    pmid = np.asarray(pmid)
    fdt = conf.float_dtype
    if val is None:
        val = default_val
    val = np.asarray(val, dtype=fdt)
    return pmid, val


def scatter(pmid, disp, conf, mesh=None, val=None, offset=0, cell_size=None):
    """``_scatter`` (``pmwd/scatter.py:33-57``) with ``_scatter_chunk`` (``:60-83``)."""
    pmid = np.asarray(pmid)
    disp = np.asarray(disp, dtype=conf.float_dtype)
    ptcl_num, spatial_ndim = pmid.shape
    fdt = conf.float_dtype

    if val is None:
        val = conf.mesh_size / conf.ptcl_num
    val = np.asarray(val, dtype=fdt)
    if mesh is None:
        mesh = np.zeros(conf.mesh_shape + val.shape[1:], dtype=fdt)
    mesh = np.array(mesh, dtype=fdt)  # copy: inputs are never modified
    if mesh.shape[spatial_ndim:] != val.shape[1:]:
        raise ValueError('channel shape mismatch: '
                         f'{mesh.shape[spatial_ndim:]} != {val.shape[1:]}')

    spatial_shape = mesh.shape[:spatial_ndim]
    chan_shape = mesh.shape[spatial_ndim:]
    flat = mesh.reshape((int(np.prod(spatial_shape)),) + chan_shape)

    for lo, hi in _chunks(ptcl_num, conf.chunk_size):
        ind, frac = enmesh(pmid[lo:hi], disp[lo:hi], conf.cell_size, conf.mesh_shape,
                           offset, cell_size, spatial_shape, False)
        lin, valid = _valid_linear(ind, spatial_shape)
        v = val[lo:hi, np.newaxis] if val.ndim != 0 else val   # scatter.py:74-75
        frac = frac.reshape(frac.shape + (1,) * len(chan_shape))
        upd = (v * frac).astype(fdt)
        upd = np.broadcast_to(upd, lin.shape + chan_shape)
        # scatter.py:80: updates applied in (particle, neighbour) order
        np.add.at(flat, lin[valid], upd[valid])
    return flat.reshape(mesh.shape)


def gather(pmid, disp, conf, mesh, val=0, offset=0, cell_size=None):
    """``_gather`` (``pmwd/gather.py:33-55``) with ``_gather_chunk`` (``:58-77``)."""
    pmid = np.asarray(pmid)
    disp = np.asarray(disp, dtype=conf.float_dtype)
    ptcl_num, spatial_ndim = pmid.shape
    fdt = conf.float_dtype

    mesh = np.asarray(mesh, dtype=fdt)
    val = np.asarray(val, dtype=fdt)
    if mesh.shape[spatial_ndim:] != val.shape[1:]:
        raise ValueError('channel shape mismatch: '
                         f'{mesh.shape[spatial_ndim:]} != {val.shape[1:]}')

    spatial_shape = mesh.shape[:spatial_ndim]
    chan_shape = mesh.shape[spatial_ndim:]
    flat = mesh.reshape((int(np.prod(spatial_shape)),) + chan_shape)
    out = np.empty((ptcl_num,) + chan_shape, dtype=fdt)

    for lo, hi in _chunks(ptcl_num, conf.chunk_size):
        ind, frac = enmesh(pmid[lo:hi], disp[lo:hi], conf.cell_size, conf.mesh_shape,
                           offset, cell_size, spatial_shape, False)
        lin, valid = _valid_linear(ind, spatial_shape)
        frac = frac.reshape(frac.shape + (1,) * len(chan_shape))
        g = flat[lin] * valid.reshape(valid.shape + (1,) * len(chan_shape))
        v = val[lo:hi] if val.ndim != 0 else val
        # gather.py:75; neighbour sum taken sequentially n = 0 .. 2**dim-1
        acc = np.zeros((hi - lo,) + chan_shape, dtype=fdt)
        for n in range(lin.shape[1]):
            acc = acc + (g[:, n] * frac[:, n]).astype(fdt)
        out[lo:hi] = v + acc
    return out


def _adj_common(pmid, disp, conf, meshlike, offset, cell_size):
    spatial_ndim = pmid.shape[1]
    spatial_shape = meshlike.shape[:spatial_ndim]
    chan_shape = meshlike.shape[spatial_ndim:]
    return spatial_ndim, spatial_shape, chan_shape


def scatter_adj(pmid, disp, conf, mesh_cot, val=None, offset=0, cell_size=None):
    """``_scatter_bwd`` / ``_scatter_chunk_adj`` (``pmwd/scatter.py:86-148``).

    Returns ``(disp_cot, val_cot)``; ``mesh_cot`` passes through unchanged and
    ``val_cot`` has the per-particle shape ``(ptcl_num,) + chan_shape``.
    """
    pmid = np.asarray(pmid)
    disp = np.asarray(disp, dtype=conf.float_dtype)
    fdt = conf.float_dtype
    ptcl_num = len(pmid)
    if val is None:
        val = conf.mesh_size / conf.ptcl_num
    val = np.asarray(val, dtype=fdt)
    mesh_cot = np.asarray(mesh_cot, dtype=fdt)
    spatial_ndim, spatial_shape, chan_shape = _adj_common(pmid, disp, conf, mesh_cot,
                                                          offset, cell_size)
    chan_axis = tuple(range(-len(chan_shape), 0)) if chan_shape else ()
    flat = mesh_cot.reshape((int(np.prod(spatial_shape)),) + chan_shape)
    disp_cot = np.empty((ptcl_num, spatial_ndim), dtype=fdt)
    val_cot = np.empty((ptcl_num,) + chan_shape, dtype=fdt)
    cs = fdt.type(cell_size if cell_size is not None else conf.cell_size)

    for lo, hi in _chunks(ptcl_num, conf.chunk_size):
        ind, frac, frac_grad = enmesh(pmid[lo:hi], disp[lo:hi], conf.cell_size,
                                      conf.mesh_shape, offset, cell_size,
                                      spatial_shape, True)
        lin, valid = _valid_linear(ind, spatial_shape)
        v = val[lo:hi, np.newaxis] if val.ndim != 0 else val
        vc = flat[lin] * valid.reshape(valid.shape + (1,) * len(chan_shape))  # :112
        dc = (vc * v)
        if chan_axis:
            dc = dc.sum(axis=chan_axis, dtype=fdt)                              # :114
        dc = _sum_axis1(dc[..., np.newaxis] * frac_grad)                        # :115
        disp_cot[lo:hi] = dc / cs                                               # :116
        fr = frac.reshape(frac.shape + (1,) * len(chan_shape))
        val_cot[lo:hi] = _sum_axis1(vc * fr)                                    # :118-119
    return disp_cot, val_cot


def gather_adj(pmid, disp, conf, mesh, val_cot, offset=0, cell_size=None):
    """``_gather_bwd`` / ``_gather_chunk_adj`` (``pmwd/gather.py:80-142``).

    Returns ``(disp_cot, mesh_cot)``; ``val_cot`` passes through unchanged.
    """
    pmid = np.asarray(pmid)
    disp = np.asarray(disp, dtype=conf.float_dtype)
    fdt = conf.float_dtype
    ptcl_num = len(pmid)
    mesh = np.asarray(mesh, dtype=fdt)
    val_cot = np.asarray(val_cot, dtype=fdt)
    spatial_ndim, spatial_shape, chan_shape = _adj_common(pmid, disp, conf, mesh,
                                                          offset, cell_size)
    chan_axis = tuple(range(-len(chan_shape), 0)) if chan_shape else ()
    flat = mesh.reshape((int(np.prod(spatial_shape)),) + chan_shape)
    mesh_cot = np.zeros_like(flat)
    disp_cot = np.empty((ptcl_num, spatial_ndim), dtype=fdt)
    cs = fdt.type(cell_size if cell_size is not None else conf.cell_size)

    for lo, hi in _chunks(ptcl_num, conf.chunk_size):
        ind, frac, frac_grad = enmesh(pmid[lo:hi], disp[lo:hi], conf.cell_size,
                                      conf.mesh_shape, offset, cell_size,
                                      spatial_shape, True)
        lin, valid = _valid_linear(ind, spatial_shape)
        vc = val_cot[lo:hi, np.newaxis] if val_cot.ndim != 0 else val_cot       # :101-102
        m = flat[lin] * valid.reshape(valid.shape + (1,) * len(chan_shape))     # :106
        dc = (vc * m)
        if chan_axis:
            dc = dc.sum(axis=chan_axis, dtype=fdt)                              # :108
        dc = _sum_axis1(dc[..., np.newaxis] * frac_grad)                        # :109
        disp_cot[lo:hi] = dc / cs                                               # :110
        fr = frac.reshape(frac.shape + (1,) * len(chan_shape))
        upd = np.broadcast_to((vc * fr).astype(fdt), lin.shape + chan_shape)
        np.add.at(mesh_cot, lin[valid], upd[valid])                             # :113
    return disp_cot, mesh_cot.reshape(mesh.shape)


def _sum_axis1(x):
    """Sequential sum over the neighbour axis, n = 0 .. 2**dim-1."""
    acc = np.zeros(x.shape[:1] + x.shape[2:], dtype=x.dtype)
    for n in range(x.shape[1]):
        acc = acc + x[:, n]
    return acc


def gen_grid(conf, vel=False, acc=False):
    """``Particles.gen_grid`` (``pmwd/particles.py:109-144``).  Returns
    ``(pmid, disp, vel, acc)`` with particle order = C-order ravel of the grid."""
    pmid, disp = [], []
    for sp, sm in zip(conf.ptcl_grid_shape, conf.mesh_shape):
        p1 = np.linspace(0, sm, num=sp, endpoint=False)
        p1 = np.rint(p1).astype(conf.pmid_dtype)
        pmid.append(p1)
        d1 = np.arange(sp) * sm - p1.astype(int) * sp      # exact int arithmetic
        d1 = d1 * (conf.cell_size / sp)
        disp.append(d1.astype(conf.float_dtype))
    pmid = np.stack(np.meshgrid(*pmid, indexing='ij'), axis=-1).reshape(-1, conf.dim)
    disp = np.stack(np.meshgrid(*disp, indexing='ij'), axis=-1).reshape(-1, conf.dim)
    v = np.zeros_like(disp) if vel else None
    a = np.zeros_like(disp) if acc else None
    return pmid, disp, v, a


def ptcl_pos(pmid, disp, conf, dtype=np.float64, wrap=True):
    """``Particles.pos`` (``pmwd/particles.py:184-209``)."""
    pos = pmid.astype(dtype)
    pos *= conf.cell_size
    pos += disp.astype(dtype)
    if wrap:
        pos %= np.array(conf.box_size, dtype=dtype)
    return pos
