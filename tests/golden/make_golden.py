"""Regenerates tests/golden/pm_small.npz with the float32 ORACLE (the reference itself cannot be
imported in this image: no jax / mcfit -- SURVEY.md F3).  These vectors therefore pin the
oracle + CUDA path against regressions; the reference-derived pins are the known-answer tests
in tests/test_oracle_pins.py.    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402


def build():
    conf = O.Conf(1., (8, 8, 8), mesh_shape=2, a_start=1 / 16, a_nbody_maxstep=1 / 4)
    cosmo = O.boltzmann(O.SimpleLCDM(conf), conf)
    white = O.white_noise(7, conf, real=True)
    ic = O.lpt(O.linear_modes(white, cosmo, conf), cosmo, conf)
    dens = O.scatter(ic['pmid'], ic['disp'], conf)
    acc = O.gravity(ic['pmid'], ic['disp'], cosmo.Omega_m, conf)
    final = O.nbody(dict(ic), cosmo, conf)
    rng = np.random.default_rng(3)
    cot = dict(disp=rng.standard_normal(ic['disp'].shape).astype(np.float32),
               vel=rng.standard_normal(ic['disp'].shape).astype(np.float32),
               acc=np.zeros_like(ic['disp']))
    _, pc, cc = O.nbody_adj(final, cot, cosmo, conf)
    return dict(white=white, growth=cosmo.growth, pmid=ic['pmid'], ic_disp=ic['disp'], ic_vel=ic['vel'],
                dens=dens, acc=acc, final_disp=final['disp'], final_vel=final['vel'], final_acc=final['acc'],
                cot_disp=cot['disp'], cot_vel=cot['vel'], adj_disp=pc['disp'], adj_vel=pc['vel'],
                adj_acc=pc['acc'], adj_Omega_m=np.float64(cc['Omega_m']), adj_growth=cc['growth'])


if __name__ == '__main__':
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'pm_small.npz')
    np.savez_compressed(out, **build())
    print('wrote', out, os.path.getsize(out), 'bytes')
