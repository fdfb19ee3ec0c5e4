"""Lab: is cuFFT's batched 2-D R2C / C2R over (y, z) the fastest way to do the (y, z) transforms of the
force pipeline, or do two 1-D plans (contiguous z pass + strided y pass) beat it?  1024 planes of 1024^2."""
import sys
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn((n, n, n), device='cuda')


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t2d = timed(lambda: torch.fft.rfft2(x))
s = torch.fft.rfft2(x)
t2di = timed(lambda: torch.fft.irfft2(s, s=(n, n), norm='forward'))
tz = timed(lambda: torch.fft.rfft(x, dim=2))
sz = torch.fft.rfft(x, dim=2)
ty = timed(lambda: torch.fft.fft(sz, dim=1))
tyi = timed(lambda: torch.fft.ifft(s, dim=1, norm='forward'))
sy = torch.fft.ifft(s, dim=1, norm='forward')
tzi = timed(lambda: torch.fft.irfft(sy, n=n, dim=2, norm='forward'))
print(f'n={n}: rfft2 {t2d:.2f} ms | rfft(z) {tz:.2f} + fft(y) {ty:.2f} = {tz + ty:.2f} ms || '
      f'irfft2 {t2di:.2f} ms | ifft(y) {tyi:.2f} + irfft(z) {tzi:.2f} = {tyi + tzi:.2f} ms')
