"""Build libcacto_b200.so in-tree with nvcc for sm_100a (one object per .cu, compiled in parallel) and the PyTorch custom-op shim
libcacto_b200_torch.so (csrc/torch_ops.cpp: TORCH_LIBRARY(cacto, ...) over the C ABI) with g++ against the installed torch."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libcacto_b200.so')
OUT_TORCH = os.path.join(HERE, 'libcacto_b200_torch.so')
# the system g++ on purpose (not $CXX): the image's /opt/gcc toolchain links parts of its own libstdc++ statically into a shared
# object, and iostream code of that copy crashes inside a process that already runs the system libstdc++ (torch)
CXX = os.environ.get('CACTO_B200_CXX', '/usr/bin/g++')
BUILD = os.path.join(ROOT, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-I' + os.path.join(ROOT, 'include'), '-I' + CSRC]
# bit-exact fp64 files must not contract a*b+c into FMA
PER_FILE = {'segtree.cu': ['-fmad=false'], 'rtg.cu': ['-fmad=false']}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')] + [os.path.join(ROOT, 'include', 'cacto_b200.h')]
    objs, jobs = [], []
    for f in sources():
        src, obj = os.path.join(CSRC, f), os.path.join(BUILD, f[:-3] + '.o')
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([NVCC] + ARCH + FLAGS + PER_FILE.get(f, []) + ['-c', src, '-o', obj])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        return r
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(OUT, objs):
        run([NVCC] + ARCH + ['-shared', '-o', OUT] + objs)
    build_torch_ops(force=force, run=run)
    return OUT


def build_torch_ops(force=False, run=None):
    """csrc/torch_ops.cpp -> libcacto_b200_torch.so, linked against libcacto_b200.so (rpath $ORIGIN) and libtorch."""
    import torch
    src = os.path.join(CSRC, 'torch_ops.cpp')
    hdr = os.path.join(ROOT, 'include', 'cacto_b200.h')
    if not (force or _stale(OUT_TORCH, [src, hdr, OUT])):
        return OUT_TORCH
    ti = os.path.dirname(torch.__file__)
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    cmd = [CXX, '-O2', '-std=c++17', '-fPIC', '-shared', '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           '-I' + os.path.join(ROOT, 'include'), '-I' + os.path.join(ti, 'include'), '-I' + os.path.join(ti, 'include', 'torch', 'csrc', 'api', 'include'),
           '-I' + cuda_inc, src, '-o', OUT_TORCH, '-L' + HERE, '-lcacto_b200', '-L' + os.path.join(ti, 'lib'), '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch',
           '-Wl,-rpath,$ORIGIN', '-Wl,-rpath,' + os.path.join(ti, 'lib')]
    if run is None:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('torch ops build failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
    else:
        run(cmd)
    return OUT_TORCH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
