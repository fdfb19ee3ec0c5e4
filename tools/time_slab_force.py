"""A/B timing of the slab force / force_adj pipeline variants in ONE process per rank (same box, same
particles): torchrun --nproc-per-node N tools/time_slab_force.py [n_per_gpu] [variants ...]
A variant is KEY=VALUE[,KEY=VALUE] of environment switches read at call time (PMWD_PIPE, PMWD_P2P_CE,
PMWD_P2P_CE_STREAMS)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
import pmwd_b200 as pm  # noqa: E402
from pmwd_b200 import dist as pd  # noqa: E402


def main():
    args = [a for a in sys.argv[1:]]
    n = int(args[0]) if args and args[0].isdigit() else 512
    variants = [a for a in args if '=' in a] or ['PMWD_PIPE=1', 'PMWD_PIPE=2', 'PMWD_PIPE=1', 'PMWD_PIPE=2']
    pd.init_process_group()
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    shape = B.rank_grid(n, world)
    conf = pm.Configuration(1., shape, mesh_shape=2, device=dev)
    comm = pd.SlabComm(conf)
    ptcl = pm.Particles.gen_grid(conf.replace(ptcl_grid_shape=(shape[0] // world,) + tuple(shape[1:]),
                                              mesh_shape=(conf.mesh_shape[0] // world,) + tuple(conf.mesh_shape[1:])),
                                 device=dev)
    pmid = ptcl.pmid.clone()
    pmid[:, 0] += comm.x0
    g = torch.Generator(device=dev).manual_seed(rank)
    disp = 2.0 * conf.cell_size * torch.randn(ptcl.disp.shape, device=dev, generator=g)
    pi = torch.randn(ptcl.disp.shape, device=dev, generator=g)
    acc, alpha = torch.empty_like(disp), torch.empty_like(disp)
    F = pd.SlabForce(conf, comm)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        fn(); fn()
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for v in variants:
        for kv in v.split(','):
            k, val = kv.split('=')
            os.environ[k] = val
        tf = timed(lambda: F.force(pmid, disp, 0.3, acc), 6)
        ta = timed(lambda: F.force_adj(pmid, disp, 0.3, pi, acc, alpha), 4)
        if rank == 0:
            print(f'{v:40s} force {tf:7.2f} ms   force_adj {ta:7.2f} ms   (halo {F.h_alloc} planes, {world} GPUs, '
                  f'{shape[0]}x{shape[1]}x{shape[2]} particles)', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
