"""Build libstgcn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python realtime-st-gcn_b200/csrc/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libstgcn_b200.so')
SOURCES = ['stgcn_api.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-shared', '-Xcompiler', '-fPIC', '-lcuda',
]


def _deps():
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cu', '.cuh'))]
    deps.append(os.path.join(HERE, '..', '..', 'include', 'stgcn_b200.h'))
    return deps


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    extra = os.environ.get('STGCN_NVCC_EXTRA', '').split()     # e.g. -DSTGCN_G3_DEBUG (role cycle counters)
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + \
          ['-o', LIB] + [os.path.join(HERE, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (proc.stdout, proc.stderr))
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
