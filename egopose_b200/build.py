"""Builds libegopose_b200.so (C ABI, include/egopose_b200.h) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libegopose_b200.so')
SOURCES = ['update_kernels.cu', 'rollout.cu', 'ozaki.cu', 'oz_mlp.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--shared', '-Xptxas', '-v']


def nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.exists(cand) or cand == 'nvcc'):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), 'include', 'egopose_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    extra = os.environ.get('EGP_NVCC_EXTRA', '').split()
    cmd = [nvcc()] + NVCC_FLAGS + extra + ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    # the container's CC wrapper lacks a few runtime bits; pin the host compiler when present
    if os.path.exists('/usr/bin/g++'):
        cmd += ['-ccbin', '/usr/bin/g++']
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, 'build.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if verbose:
        print(log)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + log[-4000:])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
