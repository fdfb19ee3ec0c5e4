"""Eulerian particle ownership for the slab decomposition (wired into the slab integrator's storage
re-sort, ``nbody._Store.reorder``; ``PMWD_MIGRATE=0`` keeps Lagrangian ownership).

With Lagrangian ownership the mesh halo has to cover the largest displacement (56 planes at N = 2,
88 planes of 2048^2 at N = 8 late in a run).  If a particle instead lives on the rank that owns its
CURRENT base plane, the halo only covers the drift since the last migration (8 planes), at the price of
moving the particles that crossed a slab boundary.  This module holds that exchange as plain tensor code -- rank
assignment with the kernels' own float32 cell arithmetic (``pmwd/pm_util.py:129-136``), a stable
partition, one count exchange and one variable-size all-to-all per array, and the inverse that
restores the reference's Lagrangian order -- so that it runs under gloo on the CPU
(``tests/test_dist_cpu.py``) as well as under NCCL.
"""
import torch
import torch.distributed as dist


def owner_rank(pmid_x, disp_x, conf, nranks):
    """Rank whose x-slab holds each particle's base cell: ``floor(disp / cell)`` in float32,
    added to the int16 ``pmid`` and wrapped periodically, exactly as ``enmesh`` does."""
    Mx = conf.mesh_shape[0]
    t = disp_x.to(torch.float32) / float(torch.tensor(conf.cell_size, dtype=torch.float32))
    cell = (pmid_x.to(torch.int64) + torch.floor(t).to(torch.int64)) % Mx
    return cell // (Mx // nranks)


def _counts(dest, nranks, group):
    send = torch.bincount(dest, minlength=nranks).to(torch.int64)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return send, recv


def exchange(arrays, dest, group=None):
    """Send row ``i`` of every tensor in ``arrays`` (dict name -> ``(n, ...)``) to rank
    ``dest[i]``.  Returns the received rows, ordered by source rank and, within one source, in the
    sender's order (stable), plus the ``(send, recv)`` row counts per rank."""
    nranks = dist.get_world_size(group)
    order = torch.sort(dest, stable=True).indices
    send, recv = _counts(dest, nranks, group)
    ssz, rsz = send.tolist(), recv.tolist()
    out = {}
    nrecv = int(sum(rsz))
    for name, a in arrays.items():
        # rows travel as raw bytes: one code path for int16 / int64 / float32 on every backend
        rows = a.index_select(0, order).contiguous()
        row_bytes = a.element_size()
        for d in a.shape[1:]:
            row_bytes *= int(d)
        src = rows.view(torch.uint8).reshape(rows.shape[0], row_bytes)
        got = torch.empty((nrecv, row_bytes), dtype=torch.uint8, device=a.device)
        dist.all_to_all_single(got, src, output_split_sizes=rsz, input_split_sizes=ssz, group=group)
        out[name] = got.view(a.dtype).reshape((nrecv,) + tuple(a.shape[1:]))
    return out, (send, recv)


def to_eulerian(arrays, conf, group=None):
    """Move every particle to the rank that owns its current base cell.  ``arrays`` needs
    ``'pmid'`` and ``'disp'``; all entries travel together (velocities, cotangents, the Lagrangian
    index ``'lag'`` that ``to_lagrangian`` needs)."""
    nranks = dist.get_world_size(group)
    dest = owner_rank(arrays['pmid'][:, 0], arrays['disp'][:, 0], conf, nranks)
    return exchange(arrays, dest, group)


def to_eulerian_movers(arrays, conf, group=None):
    """(First version of the integrator's exchange, kept as the reference for :func:`exchange_movers` in the
    tests: it also compacts the stayers, which the re-sort now does on the fly.)
    Same result set as :func:`to_eulerian` (order differs: the particles that stay come first, in
    their old order, then the arrivals by source rank), but only the particles that change rank travel:
    the owner comes from one fused pass (``pmwd_slab_owner``), the movers (a few per cent) are compacted and
    exchanged, the stayers are compacted with one gather per array."""
    from . import _lib
    import ctypes as C
    nranks, rank = dist.get_world_size(group), dist.get_rank(group)
    pmid, disp = arrays['pmid'], arrays['disp']
    n = pmid.shape[0]
    dev = disp.device
    Mx = conf.mesh_shape[0]
    if disp.is_cuda and pmid.dtype == torch.int16:
        owner = torch.empty(n, dtype=torch.uint8, device=dev)
        need = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pmwd_slab_owner(_lib.stream_ptr(dev), n, _lib.ptr(pmid), _lib.ptr(disp),
                                                  float(conf.cell_size), Mx, nranks, 0, Mx // nranks,
                                                  _lib.ptr(owner), _lib.ptr(need)), 'pmwd_slab_owner')
    else:
        owner = owner_rank(pmid[:, 0], disp[:, 0], conf, nranks).to(torch.uint8)
    stay = owner == rank
    idx_stay = stay.nonzero(as_tuple=False).squeeze(1)
    idx_move = (~stay).nonzero(as_tuple=False).squeeze(1)
    dest = owner.index_select(0, idx_move).to(torch.int64)
    order = torch.sort(dest, stable=True).indices
    idx_move = idx_move.index_select(0, order)                  # movers grouped by destination rank
    send, recv = _counts(dest, nranks, group)
    ssz, rsz = send.tolist(), recv.tolist()
    nstay, nrecv = int(idx_stay.numel()), int(sum(rsz))
    out = {}
    for name, a in arrays.items():
        res = torch.empty((nstay + nrecv,) + tuple(a.shape[1:]), dtype=a.dtype, device=a.device)
        torch.index_select(a, 0, idx_stay, out=res[:nstay])    # stayers: one gather straight into place
        rows = a.index_select(0, idx_move).contiguous()
        row_bytes = a.element_size()
        for d in a.shape[1:]:
            row_bytes *= int(d)
        src = rows.view(torch.uint8).reshape(rows.shape[0], row_bytes)
        got = res[nstay:].view(torch.uint8).reshape(nrecv, row_bytes)      # arrivals land behind the stayers
        dist.all_to_all_single(got, src, output_split_sizes=rsz, input_split_sizes=ssz, group=group)
        out[name] = res
    return out, (send, recv)


def exchange_movers(arrays, conf, group=None):
    """What the integrator's re-sort uses: the owner of every local particle (uint8, one fused pass,
    ``pmwd_slab_owner``) and the ARRIVALS only -- the particles that stay are neither copied nor compacted;
    the caller drops the rows with ``owner != rank`` (``pmwd_cell_sort_perm2`` sorts them to the end).
    Returns ``(owner, arrivals, nmove)``: ``arrivals`` maps every name to the received rows (by source rank),
    ``nmove`` is the number of local rows that left."""
    from . import _lib
    nranks, rank = dist.get_world_size(group), dist.get_rank(group)
    pmid, disp = arrays['pmid'], arrays['disp']
    n = pmid.shape[0]
    dev = disp.device
    Mx = conf.mesh_shape[0]
    if disp.is_cuda and pmid.dtype == torch.int16:
        owner = torch.empty(n, dtype=torch.uint8, device=dev)
        need = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pmwd_slab_owner(_lib.stream_ptr(dev), n, _lib.ptr(pmid), _lib.ptr(disp),
                                                  float(conf.cell_size), Mx, nranks, 0, Mx // nranks,
                                                  _lib.ptr(owner), _lib.ptr(need)), 'pmwd_slab_owner')
    else:
        owner = owner_rank(pmid[:, 0], disp[:, 0], conf, nranks).to(torch.uint8)
    idx_move = (owner != rank).nonzero(as_tuple=False).squeeze(1)
    dest = owner.index_select(0, idx_move).to(torch.int64)
    idx_move = idx_move.index_select(0, torch.sort(dest, stable=True).indices)     # grouped by destination rank
    send, recv = _counts(dest, nranks, group)
    ssz, rsz = send.tolist(), recv.tolist()
    nrecv = int(sum(rsz))
    arrivals = {}
    for name, a in arrays.items():
        rows = a.index_select(0, idx_move).contiguous()
        row_bytes = a.element_size()
        for d in a.shape[1:]:
            row_bytes *= int(d)
        src = rows.view(torch.uint8).reshape(rows.shape[0], row_bytes)
        got = torch.empty((nrecv, row_bytes), dtype=torch.uint8, device=a.device)
        dist.all_to_all_single(got, src, output_split_sizes=rsz, input_split_sizes=ssz, group=group)
        arrivals[name] = got.view(a.dtype).reshape((nrecv,) + tuple(a.shape[1:]))
    return owner, arrivals, int(idx_move.numel())


def to_lagrangian(arrays, ptcl_num, group=None):
    """Inverse of any sequence of ``to_eulerian`` calls: send every particle back to the rank that
    holds its index range of the reference's particle array and sort locally by ``arrays['lag']``
    (global index, int64).  ``ptcl_num`` must be divisible by the number of ranks."""
    nranks = dist.get_world_size(group)
    per = ptcl_num // nranks
    got, _ = exchange(arrays, arrays['lag'] // per, group)
    order = torch.sort(got['lag']).indices
    return {k: v.index_select(0, order) for k, v in got.items()}
