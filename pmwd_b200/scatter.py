"""CIC scatter with the reference's signature and VJP contract
(``pmwd/scatter.py:8-150``), executed by the CUDA kernels behind ``pmwd_scatter`` /
``pmwd_scatter_adj``.  The ``custom_vjp`` becomes a ``torch.autograd.Function``.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import CicDesc, SCATTER_ATOMIC, SCATTER_DETERMINISTIC


def make_desc(conf, pmid, spatial_shape, nchan, offset, cell_size):
    """Arguments that scatter/gather hand to ``enmesh`` (``pmwd/scatter.py:69-71``)."""
    dim = pmid.shape[1]
    d = CicDesc()
    d.dim = dim
    d.pmid_bytes = pmid.element_size()
    d.ptcl_num = pmid.shape[0]
    for a in range(3):
        d.wrap_shape[a] = conf.mesh_shape[a] if a < dim else 1
        d.mesh_shape[a] = int(spatial_shape[a]) if a < dim else 1
    d.nchan = nchan
    d.general = 0 if cell_size is None else 1
    d.cell_size = float(conf.cell_size)
    d.cell_size2 = float(cell_size) if cell_size is not None else 0.0
    if isinstance(offset, torch.Tensor):
        offset = offset.detach().cpu().tolist()
    off = [float(offset)] * dim if not isinstance(offset, (list, tuple)) else [float(o) for o in offset]
    if len(off) != dim:
        raise ValueError(f'offset must be a scalar or have {dim} entries')
    for a in range(3):
        d.offset[a] = off[a] if a < dim else 0.0
    return d


def _prep_ptcl(pmid, disp, conf):
    _lib.require_cuda(pmid, disp)
    if pmid.ndim != 2 or pmid.shape[1] not in (1, 2, 3):
        raise ValueError(f'pmid must have shape (ptcl_num, dim), dim in 1..3; got {tuple(pmid.shape)}')
    if pmid.dtype not in (torch.int8, torch.int16, torch.int32):
        raise ValueError(f'unsupported pmid dtype {pmid.dtype}')
    pmid = pmid.contiguous()
    disp = disp.to(conf.float_dtype).contiguous()
    if disp.shape != pmid.shape:
        raise ValueError(f'disp shape {tuple(disp.shape)} != pmid shape {tuple(pmid.shape)}')
    return pmid, disp


def _prep_val(val, conf, device):
    """Returns ``(tensor or None, scalar, chan_shape, is_0d_tensor)``."""
    if isinstance(val, torch.Tensor):
        val = val.to(device=device, dtype=conf.float_dtype)
        if val.ndim == 0:
            return None, float(val.item()), (), True
        return val.contiguous(), 0.0, tuple(val.shape[1:]), False
    return None, float(val), (), False


class _Scatter(torch.autograd.Function):
    """``_scatter`` + ``_scatter_fwd/_bwd`` (``pmwd/scatter.py:33-150``)."""

    @staticmethod
    def forward(ctx, pmid, disp, conf, mesh, val, offset, cell_size):
        pmid, disp = _prep_ptcl(pmid, disp, conf)
        ptcl_num, spatial_ndim = pmid.shape
        dev = disp.device

        if val is None:
            val = conf.mesh_size / conf.ptcl_num                    # scatter.py:37-39
        val_t, val_s, chan_shape, val_0d = _prep_val(val, conf, dev)

        if mesh is None:
            out = torch.zeros(conf.mesh_shape + chan_shape, dtype=conf.float_dtype, device=dev)
        else:
            mesh = torch.as_tensor(mesh, dtype=conf.float_dtype, device=dev)
            out = mesh.clone(memory_format=torch.contiguous_format)  # inputs are never modified
        if tuple(out.shape[spatial_ndim:]) != chan_shape:            # scatter.py:45-47
            raise ValueError('channel shape mismatch: '
                             f'{tuple(out.shape[spatial_ndim:])} != {chan_shape}')
        if val_t is not None and val_t.shape[0] != ptcl_num:
            raise ValueError('val must have one row per particle')

        nchan = math.prod(chan_shape)
        desc = make_desc(conf, pmid, out.shape[:spatial_ndim], nchan, offset, cell_size)
        mode = SCATTER_ATOMIC
        scratch, scratch_bytes = None, 0
        if conf.scatter_mode == 'deterministic':
            scratch_bytes = _lib.lib().pmwd_scatter_scratch_bytes(C.byref(desc), SCATTER_DETERMINISTIC)
            eligible = (scratch_bytes > 0 and nchan == 1 and spatial_ndim == 3
                        and cell_size is None and pmid.dtype == torch.int16
                        and tuple(out.shape[:3]) == tuple(conf.mesh_shape)
                        and all(o == 0 for o in desc.offset))
            if eligible:
                mode = SCATTER_DETERMINISTIC
                scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
            else:
                scratch_bytes = 0
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pmwd_scatter(
                _lib.stream_ptr(dev), C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp),
                _lib.ptr(val_t), val_s, _lib.ptr(out), mode, _lib.ptr(scratch), scratch_bytes),
                'pmwd_scatter')

        ctx.save_for_backward(pmid, disp, val_t)
        ctx.meta = (conf, val_s, val_0d, offset, cell_size, chan_shape, spatial_ndim,
                    tuple(out.shape[:spatial_ndim]))
        ctx.mesh_given = mesh is not None
        return out

    @staticmethod
    def backward(ctx, mesh_cot):
        pmid, disp, val_t = ctx.saved_tensors
        conf, val_s, val_0d, offset, cell_size, chan_shape, ndim, spatial_shape = ctx.meta
        dev = disp.device
        mesh_cot = mesh_cot.to(conf.float_dtype).contiguous()
        nchan = math.prod(chan_shape)
        desc = make_desc(conf, pmid, spatial_shape, nchan, offset, cell_size)
        disp_cot = torch.empty_like(disp)
        need_val = ctx.needs_input_grad[4]
        val_cot = (torch.empty((pmid.shape[0],) + chan_shape, dtype=conf.float_dtype, device=dev)
                   if need_val else None)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pmwd_scatter_adj(
                _lib.stream_ptr(dev), C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp),
                _lib.ptr(mesh_cot), _lib.ptr(val_t), val_s, _lib.ptr(disp_cot), _lib.ptr(val_cot)),
                'pmwd_scatter_adj')
        if need_val and val_0d:
            val_cot = val_cot.sum()
        # scatter.py:148: (None, disp_cot, None, mesh_cot, val_cot, None, None)
        return (None, disp_cot, None, mesh_cot if ctx.mesh_given else None, val_cot, None, None)


def scatter(ptcl, conf, mesh=None, val=None, offset=0, cell_size=None):
    """Scatter particle values to mesh multilinearly in n-D (``pmwd/scatter.py:8-30``)."""
    return _Scatter.apply(ptcl.pmid, ptcl.disp, conf, mesh, val, offset, cell_size)


def _scatter(pmid, disp, conf, mesh, val, offset, cell_size):
    return _Scatter.apply(pmid, disp, conf, mesh, val, offset, cell_size)
