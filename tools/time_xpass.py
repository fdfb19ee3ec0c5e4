"""Time the fused x-pass kernels alone (pmwd_xpass_force / pmwd_xpass_force_adj), CUDA events.
usage: time_xpass.py n [ny_local]   (mesh n^3; ny_local < n times the y-slab form of the multi-GPU path,
e.g. `2048 256` = one of 8 GPUs at 2048^3; PMWD_XPASS16_2048=1 selects the two-CTA cluster kernel)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from pmwd_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nyl = int(sys.argv[2]) if len(sys.argv) > 2 else n
reps = 5
lib = _lib.lib()
shape = (n, n, n)
nzc = n // 2 + 1
g = torch.Generator(device='cuda').manual_seed(0)
arrs = [torch.view_as_complex(torch.randn((n, nyl, nzc, 2), device='cuda', generator=g)) for _ in range(4)]
st = _lib.stream_ptr()
shp = _lib.shape_arr(shape)
out3 = (C.c_void_p * 3)(*[t.data_ptr() for t in arrs[1:]])
in3 = (C.c_void_p * 3)(*[t.data_ptr() for t in arrs[1:]])


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


fwd = timeit(lambda: _lib.check(lib.pmwd_xpass_force(st, shp, 0, nyl, 1.0, 0.5, _lib.ptr(arrs[0]), out3), 'fwd'))
adj = timeit(lambda: _lib.check(lib.pmwd_xpass_force_adj(st, shp, 0, nyl, 1.0, 0.5, in3, _lib.ptr(arrs[0])), 'adj'))
gb = 4 * n * nyl * nzc * 8 / 1e9
print(f'n={n} ny_local={nyl}  forward {fwd:.3f} ms ({gb / fwd * 1e3:.0f} GB/s)   adjoint {adj:.3f} ms ({gb / adj * 1e3:.0f} GB/s)')
