"""
Builds libpints_b200.so (hand-written CUDA for sm_100a + the C ABI of
include/pints_b200.h) in-tree with nvcc.  No GPU is needed to build.

    python -m pints_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libpints_b200.so')
SOURCES = ['pb200_api.cu', 'pb200_ode.cu', 'pb200_analytic.cu',
           'pb200_samplers.cu', 'pb200_rank.cu', 'pb200_exchange.cu']
HEADERS = [os.path.join(CSRC, 'pb200_common.cuh'),
           os.path.join(os.path.dirname(HERE), 'include', 'pints_b200.h')]
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build libpints_b200.so')
    return exe


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [__file__]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a; returns the path of the library."""
    if not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc(), '-O3', '-std=c++17', '-lineinfo', '-shared', '--threads', '4',
           '-Xcompiler', '-fPIC', '-Xcompiler', '-O3', '-cudart', 'static',
           *ARCH]
    if verbose:
        cmd += ['-Xptxas', '-v']
    # development: extra -D flags for kernel variants (PB200_NVCC_FLAGS="-DPB200_X=1 ...")
    cmd += os.environ.get('PB200_NVCC_FLAGS', '').split()
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libpints_b200.so')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
