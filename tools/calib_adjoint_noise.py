"""Calibration of tests/test_gpu_gravity.py::test_nbody_adjoint_config1_size_vs_oracle: how far does the float32 ORACLE
sit from the float64 one on the cosmology / particle cotangents of BASELINE config 1 (64^3, 10 steps), over 4 draws
of a 1e-6-cell perturbation of the ICs?  CPU only, ~6 minutes.  Result: profiles/r02_adjoint_noise_calibration.txt"""
import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import oracle as O
n = 64
oconf = O.Conf(1., (n,)*3, mesh_shape=2, a_nbody_maxstep=0.1)
o64 = O.Conf(1., (n,)*3, mesh_shape=2, float_dtype=np.float64, a_nbody_maxstep=0.1)
ocosmo = O.boltzmann(O.SimpleLCDM(oconf), oconf)
omodes = O.linear_modes(O.white_noise(0, oconf), ocosmo, oconf)
ic = O.lpt(omodes, ocosmo, oconf)
rng = np.random.default_rng(5)
w_disp = rng.standard_normal(ic['disp'].shape).astype(np.float32)
w_vel = rng.standard_normal(ic['disp'].shape).astype(np.float32)
t0 = time.time()
ic64 = dict(pmid=ic['pmid'], disp=ic['disp'].astype(np.float64), vel=ic['vel'].astype(np.float64))
final = O.nbody(ic64, ocosmo, o64)
cot = dict(disp=w_disp.astype(np.float64), vel=w_vel.astype(np.float64), acc=np.zeros_like(ic64['disp']))
_, pc, cc = O.nbody_adj(final, cot, ocosmo, o64)
print('f64 done', time.time() - t0, cc['Omega_m'], flush=True)
prng = np.random.default_rng(77)
for k in range(4):
    ic32 = dict(ic)
    if k:
        ic32['disp'] = (ic['disp'] + 1e-6 * oconf.cell_size * prng.standard_normal(ic['disp'].shape)).astype(np.float32)
    f32 = O.nbody(ic32, ocosmo, oconf)
    _, pc32, cc32 = O.nbody_adj(f32, dict(disp=w_disp, vel=w_vel, acc=np.zeros_like(w_disp)), ocosmo, oconf)
    e = np.abs(pc32['disp'] - pc['disp']) / np.sqrt(np.mean(pc['disp']**2))
    print(k, 'Om rel', abs(cc32['Omega_m'] / cc['Omega_m'] - 1), 'disp cot err/rms: median', np.median(e), 'p99', np.quantile(e, 0.99), 'rms', np.sqrt(np.mean(e**2)), time.time() - t0, flush=True)
