"""In-tree build of ``libpmwd_b200.so`` (hand-written sm_100a kernels + C ABI).

``python -m pmwd_b200.build`` or ``pmwd_b200.build.build_lib()``; nvcc cross-compiles
without a GPU.  The shared library is git-ignored but travels with the repo snapshot
to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libpmwd_b200.so')
BUILD = os.path.join(HERE, 'csrc', 'build')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
# --fmad=false: the kernels promise IEEE mul-then-add in the reference's operation order
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '--fmad=false',
         '-std=c++17', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2'] + os.environ.get('PMWD_NVCC_EXTRA', '').split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hdrs.append(os.path.join(os.path.dirname(HERE), 'include', 'pmwd_b200.h'))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    hdrs = _deps()
    objs = []
    jobs = []
    for s in srcs:
        o = os.path.join(BUILD, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {s}:\n{r.stdout}\n{r.stderr}')
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(compile_one, jobs):
                if verbose and out:
                    print(out, file=sys.stderr)
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + [
            '-gencode', 'arch=compute_100a,code=sm_100a', '-lcufft',
            '-Xlinker', '-rpath,/usr/local/cuda/lib64']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    print(build_lib(force='--force' in sys.argv, verbose='-v' in sys.argv))
