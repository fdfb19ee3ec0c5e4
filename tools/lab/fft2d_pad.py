"""Lab: do 128-byte aligned rows of the half-spectrum (row length padded from nz/2+1 = 513 to 528 complex64) make
cuFFT's batched 2-D R2C / C2R over (y, z) faster?  usage: fft2d_pad.py [n] [pads...]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pmwd_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pads = [int(c) for c in sys.argv[2:]] or [0, n // 2 + 1, n // 2 + 8, n // 2 + 16, n // 2 + 32]
lib = _lib.lib()
x = torch.randn((n, n, n), device='cuda')
back = torch.empty_like(x)
shape = _lib.shape_arr((n, n, n))
st = _lib.stream_ptr()


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


for pad in pads:
    row = pad if pad else n // 2 + 1
    spec = torch.zeros((n, n, row), dtype=torch.complex64, device='cuda')
    h = C.c_void_p()
    _lib.check(lib.pmwd_ctx_create(C.byref(h), torch.cuda.current_device()), 'create')
    _lib.check(lib.pmwd_ctx_set_fft2d_chunk(h, 0), 'chunk')
    _lib.check(lib.pmwd_ctx_set_fft2d_pad(h, pad), 'pad')
    _lib.check(lib.pmwd_ctx_reserve(h, 3, shape), 'reserve')
    tf = timed(lambda: _lib.check(lib.pmwd_fft2d_r2c(h, st, shape, _lib.ptr(x), _lib.ptr(spec)), 'r2c'))
    ref = torch.fft.rfft2(x[:2])
    err_f = float((spec[:2, :, :n // 2 + 1] - ref).abs().max() / ref.abs().max())
    ti = timed(lambda: _lib.check(lib.pmwd_fft2d_c2r(h, st, shape, _lib.ptr(spec), _lib.ptr(back)), 'c2r'))
    _lib.check(lib.pmwd_fft2d_r2c(h, st, shape, _lib.ptr(x), _lib.ptr(spec)), 'r2c')
    _lib.check(lib.pmwd_fft2d_c2r(h, st, shape, _lib.ptr(spec), _lib.ptr(back)), 'c2r')
    torch.cuda.synchronize()
    err = float((back / (n * n) - x).abs().max())
    print(f'n={n} row={row:4d} ({row * 8} B): R2C {tf:6.3f} ms  C2R {ti:6.3f} ms  fwd rel err {err_f:.1e}  '
          f'round-trip max err {err:.2e}', flush=True)
    lib.pmwd_ctx_destroy(h)
    del spec
