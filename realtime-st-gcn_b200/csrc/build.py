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


def build(force=False, verbose=False, out=None, extra_flags=()):
    """``out``/``extra_flags``: build a variant next to the library (e.g. the measurement build
    ``libstgcn_b200_dbg.so`` with ``-DSTGCN_DEBUG_BUILD``, selected at run time with ``STGCN_LIB``)."""
    if out is None and not force and not stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    extra = os.environ.get('STGCN_NVCC_EXTRA', '').split() + list(extra_flags)
    out = out or LIB
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + \
          ['-o', out] + [os.path.join(HERE, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (proc.stdout, proc.stderr))
    if verbose:
        print(proc.stderr)
    return out


if __name__ == '__main__':
    if '--debug' in sys.argv:
        print(build(out=os.path.join(HERE, 'libstgcn_b200_dbg.so'), extra_flags=['-DSTGCN_DEBUG_BUILD']))
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
