"""Timeline of pmwd_b200.nbody_step_host (host arrays in and out every step): when each stage of a call ends on
the device, relative to the first call's start.  usage: PMWD_HOST_TRACE=1 python tools/time_host_step.py [n] [steps]"""
import os
import sys

os.environ['PMWD_HOST_TRACE'] = '1'
import importlib
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmwd_b200 as pm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nb = importlib.import_module('pmwd_b200.nbody')
conf = pm.Configuration(1., (n,) * 3, mesh_shape=2)
cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
with torch.no_grad():
    ic, _ = pm.lpt(pm.linear_modes(pm.white_noise(0, conf, real=True), cosmo, conf), cosmo, conf)
    a = conf.a_nbody.tolist()
    p0, _ = pm.nbody_init(a[0], ic, None, cosmo, conf)
host = {k: getattr(p0, k).cpu().pin_memory() for k in ('pmid', 'disp', 'vel', 'acc')}
del p0, ic
torch.cuda.empty_cache()
pm.nbody_step_host(a[0], a[1], host, cosmo, conf, out=host)      # warm-up (uploads acc, plans, allocations)
torch.cuda.synchronize()
m = nb._host_mirrors[str(torch.device('cuda', torch.cuda.current_device()))]
m.trace.clear()
for j in range(1, 1 + steps):
    pm.nbody_step_host(a[j], a[j + 1], host, cosmo, conf, out=host)
torch.cuda.synchronize()
t0 = m.trace[0][1]
last = 0.0
for label, ev in m.trace:
    t = t0.elapsed_time(ev)
    print(f'{label:12s} {t:9.2f} ms' + (f'   (call took {t - last:7.2f})' if label == 'vel_down' else ''))
    if label == 'start':
        last = t
