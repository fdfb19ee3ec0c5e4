"""Per-step stage times over a full run (development aid): how scatter / gather cost evolves
with clustering, with and without storage re-ordering."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmwd_b200 as pm
from pmwd_b200 import _lib
from pmwd_b200.nbody import _integrate_inplace, _force_inplace, _store_from

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=512)
ap.add_argument('--reorder-every', type=int, default=4)
ap.add_argument('--reorder-min-disp', type=float, default=1.5)
ap.add_argument('--out', default='gpurun_out/trace.json')
a = ap.parse_args()
conf = pm.Configuration(1., (a.n,) * 3, mesh_shape=2, reorder_every=a.reorder_every,
                        reorder_min_disp=a.reorder_min_disp)
cosmo = pm.boltzmann(pm.SimpleLCDM(conf), conf)
with torch.no_grad():
    ic, _ = pm.lpt(pm.linear_modes(pm.white_noise(0, conf), cosmo, conf), cosmo, conf)
    st = _store_from(ic, conf)
    _force_inplace(st.ptcl, cosmo, conf)
    al = conf.a_nbody.tolist()
    _lib.profile_enable(True); _lib.profile_read()
    rows = []
    for i in range(len(al) - 1):
        _integrate_inplace(al[i], al[i + 1], st.ptcl, cosmo, conf)
        st.maybe_reorder()
        r = _lib.profile_read()
        row = {k: round(v[0], 2) for k, v in r.items() if v[1]}
        row['a'] = round(al[i + 1], 3)
        row['maxdisp_cells'] = round(float(st.arrays['disp'].abs().max()) / conf.cell_size, 1)
        rows.append(row)
        print(i, row, flush=True)
json.dump(rows, open(a.out, 'w'))
