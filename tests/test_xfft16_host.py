"""CPU emulation of the register-resident FFT used by csrc/xpass16.cu.

``tests/host/xfft16_emul.cc`` includes the same ``xfft16.cuh`` the kernels use and runs every
barrier-separated phase thread by thread against a naive float64 DFT (index arithmetic, exchange
buffer coverage / collisions, shared-memory bank conflicts).  No GPU needed.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('g++') is None, reason='g++ not available')
def test_register_fft_emulation(tmp_path):
    exe = tmp_path / 'xfft16_emul'
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    subprocess.run(['g++', '-O2', '-std=c++17', '-I' + cuda_inc, '-I' + os.path.join(ROOT, 'pmwd_b200', 'csrc'),
                    os.path.join(ROOT, 'tests', 'host', 'xfft16_emul.cc'), '-o', str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'OK' in r.stdout and 'bank conflicts (64-bit half-warp model): 0' in r.stdout
