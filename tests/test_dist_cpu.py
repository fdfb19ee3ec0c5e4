"""world_size-2 and -4 gloo tests (CPU) of the slab decomposition's host logic: the distributed FFT
(2-D local transforms + one all-to-all + 1-D transform) against numpy's rfftn/irfftn of the
gathered field, the neighbour halo reduce / fill, and the particle partition."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from pmwd_b200.configuration import Configuration
    from pmwd_b200.dist import SlabComm
    dist.init_process_group('gloo')
    r, P = dist.get_rank(), dist.get_world_size()
    conf = Configuration(1., (8, 4, 6), mesh_shape=2, device='cpu')
    comm = SlabComm(conf)
    Mx, My, Mz = conf.mesh_shape
    assert (comm.mx, comm.my) == (Mx // P, My // P) and comm.x0 == r * comm.mx
    sl = comm.local_slice()
    assert sl.stop - sl.start == conf.ptcl_num // P and sl.start == r * conf.ptcl_num // P

    full = np.random.default_rng(0).standard_normal(conf.mesh_shape).astype(np.float32)
    slab = torch.from_numpy(full[comm.x0:comm.x0 + comm.mx].copy())
    spec = comm.rfftn(slab)                                # [Mx][my][Mz/2+1]
    ref = np.fft.rfftn(full.astype(np.float64))
    got = spec.numpy()
    np.testing.assert_allclose(got, ref[:, comm.y0:comm.y0 + comm.my], rtol=0, atol=2e-4 * np.abs(ref).max())
    back = comm.irfftn(spec, My, Mz) / full.size
    np.testing.assert_allclose(back.numpy(), full[comm.x0:comm.x0 + comm.mx], rtol=0, atol=1e-5)

    # halo reduce: every rank deposits ones into its extended slab -> owned planes get the
    # neighbours' halo contributions;  halo fill: halos equal the neighbours' owned planes
    h = 3
    ext = torch.zeros(comm.mx + 2 * h, My, Mz)
    ext += (r + 1)
    own = comm.halo_reduce(ext.clone(), h)
    left, right = (r - 1) %% P + 1, (r + 1) %% P + 1
    exp = torch.full((comm.mx, My, Mz), float(r + 1))
    exp[:h] += left
    exp[comm.mx - h:] += right
    assert torch.equal(own, exp)
    planes = torch.arange(comm.x0, comm.x0 + comm.mx, dtype=torch.float32).reshape(-1, 1, 1).expand(comm.mx, My, Mz)
    ext3 = torch.zeros(3, comm.mx + 2 * h, My, Mz)
    ext3[:, h:h + comm.mx] = planes
    comm.halo_fill(ext3, h)
    want = (torch.arange(comm.x0 - h, comm.x0 + comm.mx + h) %% Mx).float().reshape(-1, 1, 1).expand(-1, My, Mz)
    assert torch.equal(ext3[1], want)
    assert comm.allreduce_max(torch.tensor(float(r))) == P - 1
    # transposed-layout wavevectors and the parity-mode white noise: rank slices of the global objects
    import math
    from pmwd_b200.dist import _kvec_T, white_noise_slab
    kT = _kvec_T(conf, comm, conf.mesh_shape, conf.cell_size, 'cpu')
    period = 2 * math.pi / conf.cell_size
    np.testing.assert_array_equal(kT[0].ravel().numpy(), (np.fft.fftfreq(Mx) * period).astype(np.float32))
    np.testing.assert_array_equal(kT[1].ravel().numpy(),
                                  (np.fft.fftfreq(My) * period).astype(np.float32)[comm.y0:comm.y0 + comm.my])
    np.testing.assert_array_equal(kT[2].ravel().numpy(), (np.fft.rfftfreq(Mz) * period).astype(np.float32))
    wn = white_noise_slab(3, conf, comm, 'cpu')
    full_wn = np.random.default_rng(3).standard_normal(conf.ptcl_grid_shape, dtype=np.float32)
    np.testing.assert_array_equal(wn.numpy(), full_wn[comm.px0:comm.px0 + comm.pnx])
    assert comm.pnx == conf.ptcl_grid_shape[0] // P and comm.px0 == r * comm.pnx
    # Eulerian ownership (groundwork, pmwd_b200/migrate.py): after to_eulerian every particle sits on the
    # rank that owns its base cell and nothing is lost; to_lagrangian restores the reference order exactly
    from pmwd_b200 import migrate
    import oracle as O
    oconf = O.Conf(1., conf.ptcl_grid_shape, mesh_shape=2)
    pmid_all, disp_all, _, _ = O.gen_grid(oconf)
    rng = np.random.default_rng(11)
    disp_all = (disp_all + 3.0 * rng.standard_normal(disp_all.shape)).astype(np.float32)   # crosses slabs, wraps
    vel_all = rng.standard_normal(disp_all.shape).astype(np.float32)
    n_all = len(pmid_all)
    mine = arrs = dict(pmid=torch.from_numpy(pmid_all[sl]), disp=torch.from_numpy(disp_all[sl]),
                       vel=torch.from_numpy(vel_all[sl]), lag=torch.arange(sl.start, sl.stop))
    eul, (sent, recvd) = migrate.to_eulerian(arrs, conf)
    own = migrate.owner_rank(eul['pmid'][:, 0], eul['disp'][:, 0], conf, P)
    assert bool((own == r).all()) and int(sent.sum()) == len(mine['lag'])
    ind, _ = O.enmesh(eul['pmid'].numpy(), eul['disp'].numpy(), oconf.cell_size, oconf.mesh_shape, 0, None, None, False)
    base_x = ind[:, 0, 0]
    assert ((base_x >= comm.x0) & (base_x < comm.x0 + comm.mx)).all()        # same cell arithmetic as enmesh
    total = torch.tensor([len(eul['lag'])]); dist.all_reduce(total)
    assert int(total) == n_all
    np.testing.assert_array_equal(eul['vel'].numpy(), vel_all[eul['lag'].numpy()])      # rows travel together
    back = migrate.to_lagrangian(eul, n_all)
    for k in ('pmid', 'disp', 'vel', 'lag'):
        assert torch.equal(back[k], mine[k]), k
    # the movers-only exchange the integrator uses: same particle set per rank (stayers first), same way home
    eul2, _ = migrate.to_eulerian_movers(arrs, conf)
    assert sorted(eul2['lag'].tolist()) == sorted(eul['lag'].tolist())
    np.testing.assert_array_equal(eul2['vel'].numpy(), vel_all[eul2['lag'].numpy()])
    back2 = migrate.to_lagrangian(eul2, n_all)
    for k in ('pmid', 'disp', 'vel', 'lag'):
        assert torch.equal(back2[k], mine[k]), k
    # what the re-sort consumes: owner mask + arrivals only (the stayers are not copied)
    owner, arr, nmove = migrate.exchange_movers(arrs, conf)
    stay = owner == r
    assert nmove == int((~stay).sum()) and len(arr['lag']) == len(eul['lag']) - int(stay.sum())
    assert sorted(torch.cat([mine['lag'][stay], arr['lag']]).tolist()) == sorted(eul['lag'].tolist())
    np.testing.assert_array_equal(arr['vel'].numpy(), vel_all[arr['lag'].numpy()])
    np.testing.assert_array_equal(arr['pmid'].numpy(), pmid_all[arr['lag'].numpy()])
    print('rank', r, 'ok')
''')


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.timeout(300)
@pytest.mark.parametrize('world', [2, 4])
def test_slab_host_logic_gloo(tmp_path, world):
    """World 2: every other rank is both neighbours; world 4: distinct left / right neighbours, non-neighbour
    destinations in the particle exchange, ranks that receive nothing from some peers."""
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='1')
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('ok') == world
