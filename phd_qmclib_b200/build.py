"""Build libqmcb200.so in-tree with nvcc for sm_100a (no JIT cache: the
built library travels to the GPU box with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libqmcb200.so')
SOURCES = ['qmcb_api.cu']
DEPS = ['qmcb_api.cu', 'qmcb_kernels.cuh', 'qmcb_dev.cuh',
        'qmcb_estimators.cuh', 'qmcb_vmc.cuh',
        os.path.join('..', '..', 'include', 'qmcb200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
    '-std=c++17', '-shared', '-Xcompiler', '-fPIC,-fvisibility=hidden',
    '--fmad=true', '-Xptxas', '-v',
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS
               if os.path.exists(os.path.join(CSRC, d)))


def build(force=False, verbose=False, defines=(), out=None):
    """``defines`` / ``out``: build a tuning variant next to the product
    library (development aid, see scripts/)."""
    lib = out or LIB
    if not force and not defines and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-D' + d for d in defines] + ['-o', lib] + \
        [os.path.join(CSRC, s) for s in SOURCES] + ['-ldl']
    env = dict(os.environ)
    # the image's $CC wrapper is not a usable host compiler for nvcc
    if os.path.exists('/usr/bin/g++'):
        cmd[1:1] = ['-ccbin', '/usr/bin/g++']
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libqmcb200.so')
    if not out:
        with open(os.path.join(HERE, 'build.log'), 'w') as f:
            f.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
    return lib


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
    print(LIB)
