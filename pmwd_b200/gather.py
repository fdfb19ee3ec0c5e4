"""CIC gather with the reference's signature and VJP contract (``pmwd/gather.py:8-144``),
executed by the CUDA kernels behind ``pmwd_gather`` / ``pmwd_gather_adj``.
"""
import ctypes as C
import math

import torch

from . import _lib
from .scatter import make_desc, _prep_ptcl, _prep_val


class _Gather(torch.autograd.Function):
    """``_gather`` + ``_gather_fwd/_bwd`` (``pmwd/gather.py:33-144``)."""

    @staticmethod
    def forward(ctx, pmid, disp, conf, mesh, val, offset, cell_size):
        pmid, disp = _prep_ptcl(pmid, disp, conf)
        ptcl_num, spatial_ndim = pmid.shape
        dev = disp.device
        mesh = torch.as_tensor(mesh, dtype=conf.float_dtype, device=dev).contiguous()
        _lib.require_cuda(mesh)
        val_t, val_s, chan_shape, _ = _prep_val(val, conf, dev)
        if tuple(mesh.shape[spatial_ndim:]) != chan_shape:            # gather.py:41-43
            raise ValueError('channel shape mismatch: '
                             f'{tuple(mesh.shape[spatial_ndim:])} != {chan_shape}')
        if val_t is not None and val_t.shape[0] != ptcl_num:
            raise ValueError('val must have one row per particle')
        nchan = math.prod(chan_shape)
        desc = make_desc(conf, pmid, mesh.shape[:spatial_ndim], nchan, offset, cell_size)
        out = torch.empty((ptcl_num,) + chan_shape, dtype=conf.float_dtype, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pmwd_gather(
                _lib.stream_ptr(dev), C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp),
                _lib.ptr(mesh), _lib.ptr(val_t), val_s, _lib.ptr(out)), 'pmwd_gather')
        ctx.save_for_backward(pmid, disp, mesh)
        ctx.meta = (conf, offset, cell_size, chan_shape, spatial_ndim)
        return out

    @staticmethod
    def backward(ctx, val_cot):
        pmid, disp, mesh = ctx.saved_tensors
        conf, offset, cell_size, chan_shape, ndim = ctx.meta
        dev = disp.device
        val_cot = val_cot.to(conf.float_dtype).contiguous()
        nchan = math.prod(chan_shape)
        desc = make_desc(conf, pmid, mesh.shape[:ndim], nchan, offset, cell_size)
        disp_cot = torch.empty_like(disp)
        mesh_cot = torch.zeros_like(mesh) if ctx.needs_input_grad[3] else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pmwd_gather_adj(
                _lib.stream_ptr(dev), C.byref(desc), _lib.ptr(pmid), _lib.ptr(disp),
                _lib.ptr(mesh), _lib.ptr(val_cot), 0.0, _lib.ptr(disp_cot), _lib.ptr(mesh_cot)),
                'pmwd_gather_adj')
        # gather.py:142: (None, disp_cot, None, mesh_cot, val_cot, None, None)
        return (None, disp_cot, None, mesh_cot,
                val_cot if ctx.needs_input_grad[4] else None, None, None)


def gather(ptcl, conf, mesh, val=0, offset=0, cell_size=None):
    """Gather particle values from mesh multilinearly in n-D (``pmwd/gather.py:8-30``)."""
    return _Gather.apply(ptcl.pmid, ptcl.disp, conf, mesh, val, offset, cell_size)


def _gather(pmid, disp, conf, mesh, val, offset, cell_size):
    return _Gather.apply(pmid, disp, conf, mesh, val, offset, cell_size)
