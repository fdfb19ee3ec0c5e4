"""Lab: the (y, z) 2-D transforms of the force pipeline issued `chunk` x planes per cuFFT call (pmwd_ctx_set_fft2d_chunk):
does the second pass of a chunk find the first one's output in L2?  usage: fft2d_chunks.py [n] [chunks...]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pmwd_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
chunks = [int(c) for c in sys.argv[2:]] or [0, 2, 4, 8, 16, 32, 64]
lib = _lib.lib()
x = torch.randn((n, n, n), device='cuda')
spec = torch.empty((n, n, n // 2 + 1), dtype=torch.complex64, device='cuda')
back = torch.empty_like(x)
shape = _lib.shape_arr((n, n, n))
st = _lib.stream_ptr()
ref = None


def timed(fn, reps=6):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for c in chunks:
    h = C.c_void_p()
    _lib.check(lib.pmwd_ctx_create(C.byref(h), torch.cuda.current_device()), 'create')
    _lib.check(lib.pmwd_ctx_set_fft2d_chunk(h, c), 'chunk')
    _lib.check(lib.pmwd_ctx_reserve(h, 3, shape), 'reserve')
    _lib.check(lib.pmwd_fft2d_r2c(h, st, shape, _lib.ptr(x), _lib.ptr(spec)), 'r2c')
    torch.cuda.synchronize()
    if ref is None:
        ref = spec.clone()
        same = True
    else:
        same = bool(torch.equal(torch.view_as_real(spec), torch.view_as_real(ref)))
    tf = timed(lambda: _lib.check(lib.pmwd_fft2d_r2c(h, st, shape, _lib.ptr(x), _lib.ptr(spec)), 'r2c'))
    # C2R clobbers its input: time it on whatever is there (data independent), check once on a fresh spectrum
    ti = timed(lambda: _lib.check(lib.pmwd_fft2d_c2r(h, st, shape, _lib.ptr(spec), _lib.ptr(back)), 'c2r'))
    spec.copy_(ref)
    _lib.check(lib.pmwd_fft2d_c2r(h, st, shape, _lib.ptr(spec), _lib.ptr(back)), 'c2r')
    torch.cuda.synchronize()
    err = float((back / (n * n) - x).abs().max())
    print(f'n={n} chunk={c:3d}: R2C {tf:6.3f} ms  C2R {ti:6.3f} ms  forward identical to chunk 0: {same}  '
          f'round-trip max err {err:.2e}', flush=True)
    lib.pmwd_ctx_destroy(h)
